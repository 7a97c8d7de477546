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
//     - rows are split into kc = 64*T >= k classes (position of the row inside its 64-row tile, times
//       the CTA's group blockIdx % T); per query the best normalised score of each of the CTA's 64
//       classes is tracked in REGISTERS with one FMNMX per score.  The minimum over all kc classes
//       certifies kc >= k DISTINCT rows at least that good, so any row below it (minus a 1e-6
//       safety margin) cannot be in the top-k.  Class maxima are merged across CTAs through a
//       global array (atomicMax), so the bound reflects every row the whole grid has seen.
//     - a score is tested with one multiply + compare; only survivors (a fraction ~k ln(n)/n) are
//       appended, as raw (row, score) pairs, to the (CTA, query) survivor buffer in global memory.
//       The reference's distance formula (src/metrics.rs:157-164) runs once per survivor at the
//       end of the kernel, a warp per query.
//     - the first kTcLearnTiles tiles of a CTA only feed the class maxima ("learn") and are scanned
//       again at the end, when the bound is mature.
//     - at the end each warp bitonic-sorts its queries' survivors and writes the best k_pad.
//
// * clusters: with more than 128 queries in a launch, CTAs run as clusters of C = 2; the CTAs of a
//   cluster hold different 128-query groups in their TMEM and walk the SAME store tiles.  Every ring
//   stage is fetched from HBM once: CTA r copies the r-th 1/C of the stage with
//   cp.async.bulk ... .multicast::cluster into the shared memory of all C CTAs (each CTA's full
//   barrier counts the whole stage).  A slot is refilled only when every CTA of the cluster has
//   released it: each MMA thread's tcgen05.commit is multicast to the slot's empty barrier in all C
//   CTAs (barrier count C).  This halves the HBM bytes per query and moves the kernel from
//   HBM-bound towards tensor-bound.
//
// * pairs (MODE 2, `cta_group::2`): the two CTAs of a cluster form one M = 256 MMA.  Each CTA still
//   owns 128 queries (its TMEM lanes) but loads only HALF of every store tile (rows 0-31 or 32-63,
//   12 x 4 KiB copies into its own shared memory): tcgen05.mma.cta_group::2 takes the first N/2
//   rows of B from the leader's shared memory and the rest from its peer's.  The bytes an SM
//   ingests per flop halve -- the L2->SM path is what limits the multicast variant.  CTA 1 forwards
//   "my half has landed" to the leader with a remote mbarrier arrive; the leader issues all MMAs and
//   its tcgen05.commit.cta_group::2 ... multicast releases ring slots and publishes accumulators in
//   both CTAs.
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
constexpr u32 kTcStageKB = 12;   // max 64-column blocks per ring stage: one cp.async.bulk each (a single
                                 // thread issuing 8 KiB copies cannot keep an SM fed: ~300 cycles each)
constexpr u32 kTcStageBytes = kTcStageKB * kTileBytes;
constexpr u32 kTcLearnTiles = 8; // tiles scanned bound-only first, then again at the end
constexpr u32 kTcSurvivorCap = 1024;  // survivor slots per (CTA, query) for top_k <= 512; 2 x k_pad above (TensorScanArgs::surv_cap)
constexpr u32 kTcNormSlots = 16; // ring of per-tile (|a|, 1/|a|) blocks brought in by TMA (the producer runs ~5 tiles ahead of the epilogue)

struct TensorScanArgs {
    const unsigned char* tiles;  // BF16 store (tile layout)
    const float* norms;          // |a| per row
    const float* inv_norms;      // 1/|a| per row (0 where |a| is 0, inf or NaN)
    const float* queries;        // [nq_valid][dim] binary32, device
    u64* cta_keys;               // unused (per-cluster candidate lists of earlier versions)
    u32* g_gm;                   // [64*T][128*C] global class maxima (ordered-float bits), 0 before the launch
    u64* survivors;              // [128*C][n_clusters][2 * kTcSurvivorCap] raw (row, score) pairs
    u32* surv_count;             // [128*C][n_clusters] survivors per (query, cluster)
    float* q_mag;                // [128*C] |b| of the queries as given (written by cluster 0)
    u32 nq_valid;                // <= 128 * cluster size
    u32 n_rows, n_tiles, KB, dim, k, k_pad, row_base, stages;
    u32 stage_kb;                // 64-column blocks per ring stage (<= kTcStageKB)
    // The tensor cores are a FILTER: they see bf16-rounded queries (and, for an fp32 store, its bf16
    // shadow), and every survivor's distance is then recomputed with the reference's own arithmetic
    // (sequential binary32 fold over the stored row and the UNROUNDED query) -- results are bit-identical
    // to the exact scan's.  `margin` widens the bound by twice the worst-case gap between the two dot
    // products, in units of |a| * |b| (Cauchy-Schwarz on the element-wise rounding errors).
    float margin;
    const float4* exact_panels;  // fp32 store: the exact rows (panel layout); nullptr: `tiles` ARE the exact rows (bf16 store)
    u32 exact_dq;                // fp32 store: column quads per row
    u32 surv_cap;                // survivor slots per (cluster, query) list; the buffer holds 2 x surv_cap (compaction scratch)
    u32 metric;                  // bsr_metric: 0 cosine (scores scaled by 1/|a|), 1 dot (inv_norms -> 64 ones, stride 0)
    u32 norm_stride;             // rows between consecutive tiles' entries of norms / inv_norms (64, or 0 for BSR_DOT)
    long long* timeline;         // debug builds (-DBSR_TC_TIMELINE): [2 CTAs][kTlTiles][8] clock64 stamps
    u32 tl_start;                // first tile (position in the CTA's sequence) the timeline records
};

// Debug timeline (tools/probe_timeline.sh): CTAs 0 and 1 stamp the pipeline events of their first
// kTlTiles tiles.  Compiled out unless BSR_TC_TIMELINE is defined.
constexpr u32 kTlTiles = 96;
#ifdef BSR_TC_TIMELINE
#define BSR_TL(sq, ev)                                                                                   \
    do {                                                                                                 \
        if (a.timeline && blockIdx.x < 2 && (sq) >= a.tl_start && (sq) < a.tl_start + kTlTiles)          \
            a.timeline[((size_t)blockIdx.x * kTlTiles + ((sq) - a.tl_start)) * 8 + (ev)] = clock64();   \
    } while (0)
#else
#define BSR_TL(sq, ev) do { } while (0)
#endif

// ---- tcgen05 wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One lane of a converged warp.  The tcgen05 issue loops run with all 32 lanes (uniform control
// flow, operands in uniform registers) and predicate only the instruction itself on this: under
// `if (lane == 0)` the compiler wraps every UTCHMMA in an R2UR + ELECT/BRA "waterfall" (~60 cycles
// per MMA, measured), which made the issuing thread the bottleneck of the whole kernel.
__device__ __forceinline__ bool elect_one() {
    u32 pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- cluster helpers ----
__device__ __forceinline__ u32 cluster_ctarank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default
// semantics on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR in front of the
// arrive (~2 000 cycles per tile with survivor stores in flight, measured in the pair-mode
// epilogue).  What these arrives order is tcgen05 / TMA (async proxy) work, which the tcgen05 fences
// and the barrier's own completion already cover.
__device__ __forceinline__ void mbar_arrive_cluster(u32 local_bar, u32 rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(local_bar), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(u32 bar, u32 parity) {
    u32 ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy delivered to the same offset in every CTA of `mask`; each destination
// CTA's barrier (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(u32 dst, const void* src, u32 bytes, u32 bar, unsigned short mask,
                                                   u64 policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4, %5;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask), "l"(policy)
        : "memory");
}

// the same completion signalled to the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_multicast(u32 bar, unsigned short mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair_multicast(u32 bar, unsigned short mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
// M = 256 across a CTA pair: A from both CTAs' tensor memory, B halves from both shared memories
__device__ __forceinline__ void tc_mma_ts_pair(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
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
// the same for a CTA pair: M = 256 (128 queries per CTA), N = 64
constexpr u32 kIdescBf16M256N64 = (1u << 4) | (1u << 7) | (1u << 10) | ((kTcN >> 3) << 17) | (((2 * kTcQ) >> 4) << 24);

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

// max of three, NaN if any input is NaN (SASS FMNMX3.NAN)
__device__ __forceinline__ float fmax3_nan(float x, float y, float z) {
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x), "f"(y), "f"(z));
    return r;
}

// predicated fire-and-forget atomic max (SASS @P REDG.MAX): no branch around it
__device__ __forceinline__ void red_max_if(u32* addr, u32 val, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "@p red.global.max.u32 [%0], %1;\n\t}"
        ::"l"(addr), "r"(val), "r"((u32)pred)
        : "memory");
}

// order-preserving u32 image of a float (for atomicMax on class maxima); NaN never gets here
__device__ __forceinline__ u32 f32_ordered(float f) { return ordered_from_bits(__float_as_uint(f)); }
__device__ __forceinline__ float f32_from_ordered(u32 o) { return __uint_as_float(bits_from_ordered(o)); }

// A survivor is stored raw: (local row << 32) | score bits.  The score only got it through the filter;
// its distance is the reference's own arithmetic (src/metrics.rs:143-165) over the STORED row and the
// query as given: sequential binary32 fold, identity test, zero-norm rule, clamp -- the same device
// functions as the exact scan, so the key is bit-identical to the one the exact scan would produce.
__device__ __forceinline__ u64 pack_survivor(u32 row, float score) {
    return ((u64)row << 32) | (u64)__float_as_uint(score);
}
__device__ __forceinline__ float4 load_query_quad(const float* qp, u32 c, u32 dim, bool vec_ok) {
    if (vec_ok && c + 4 <= dim) return __ldg(reinterpret_cast<const float4*>(qp + c));
    float4 q;
    q.x = c + 0 < dim ? __ldg(qp + c + 0) : 0.0f;
    q.y = c + 1 < dim ? __ldg(qp + c + 1) : 0.0f;
    q.z = c + 2 < dim ? __ldg(qp + c + 2) : 0.0f;
    q.w = c + 3 < dim ? __ldg(qp + c + 3) : 0.0f;
    return q;
}
__device__ __noinline__ u64 survivor_key(const TensorScanArgs& a, u64 raw, float na, const float* qp, float mag_b) {
    const u32 row = (u32)(raw >> 32);
    const bool vec_ok = (a.dim % 4 == 0) && ((reinterpret_cast<size_t>(qp) & 15) == 0);
    float acc = 0.0f;
    bool identical = true;
    // The row is read eight 16-byte pieces at a time, all eight loads in flight before the first is
    // folded: a survivor's row is cold (DRAM latency), and the fold itself must stay sequential.
    if (a.exact_panels) {  // fp32 store: panel layout, column quad j of row r at panels[((r/32)*dq + j)*32 + r%32]
        const float4* p = a.exact_panels + (size_t)(row >> 5) * a.exact_dq * 32 + (row & 31u);
        for (u32 j0 = 0; j0 < a.exact_dq; j0 += 8) {
            float4 v[8];
#pragma unroll
            for (u32 u = 0; u < 8; u++)
                v[u] = j0 + u < a.exact_dq ? p[(size_t)(j0 + u) * 32] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (u32 u = 0; u < 8; u++) {
                if (j0 + u < a.exact_dq) {
                    const float4 q = load_query_quad(qp, 4 * (j0 + u), a.dim, vec_ok);
                    acc = dot4_seq(acc, v[u], q);
                    identical = identical && !differs4(v[u], q);
                }
            }
        }
    } else {               // bf16 store: the tiles are the rows; 8 chunks = one 64-column block
        for (u32 kb = 0; kb < a.KB; kb++) {
            uint4 v[8];
#pragma unroll
            for (u32 u = 0; u < 8; u++) v[u] = *reinterpret_cast<const uint4*>(a.tiles + bf16_chunk_offset(row, kb, u, a.KB));
#pragma unroll
            for (u32 u = 0; u < 8; u++) {
                const u32 c = kb * 64 + u * 8;
                const float4 q0 = load_query_quad(qp, c, a.dim, vec_ok), q1 = load_query_quad(qp, c + 4, a.dim, vec_ok);
                acc = dot8_seq_bf16(acc, v[u], q0, q1);
                identical = identical && !differs8_bf16(v[u], q0, q1);
            }
        }
    }
    return make_key(finish_distance(acc, na, mag_b, identical, a.metric), a.row_base + row);
}
// Overflowing survivor buffer (adversarial row order only): keep, exactly, the k best of the cap
// entries.  Out of line: it must not bloat the score loop.  Returns the exact k-th best key.
__device__ __noinline__ u64 compact_survivors(const TensorScanArgs& a, u64* buf, u32 n, const float* qp,
                                              float mag_b, u64* keys_tmp) {
    const u32 k = a.k;
    for (u32 i = 0; i < n; i++) keys_tmp[i] = survivor_key(a, buf[i], __ldg(a.norms + (u32)(buf[i] >> 32)), qp, mag_b);
    for (u32 x = 0; x < k; x++) {  // selection of the k smallest keys, entries move with their keys
        u32 best = x;
        for (u32 y = x + 1; y < n; y++)
            if (keys_tmp[y] < keys_tmp[best]) best = y;
        const u64 tk = keys_tmp[x], tb = buf[x];
        keys_tmp[x] = keys_tmp[best];
        buf[x] = buf[best];
        keys_tmp[best] = tk;
        buf[best] = tb;
    }
    return keys_tmp[k - 1];
}

// MODE 0: one CTA per 128 queries.  MODE 1: clusters of 2, store tiles multicast.  MODE 2: CTA pairs,
// cta_group::2 (M = 256), each CTA loads half of every tile.
template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) scan_topk_bf16_tensor_kernel(const TensorScanArgs a) {
    constexpr int C = MODE == 0 ? 1 : 2;
    constexpr bool PAIR = MODE == 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 S = a.stages, KB = a.KB, SKB = a.stage_kb;
    // bytes of one 64-column block in MY shared memory: all 64 rows, or my 32 rows in pair mode
    constexpr u32 kBlk = PAIR ? kTileBytes / 2 : kTileBytes;
    const u32 blk_bytes = kBlk;
    const u32 stage_bytes = SKB * blk_bytes;
    // [ring: S stages (1024-byte aligned: swizzle atoms) | norm ring: 16 x (64 |a| + 64 1/|a|) |
    //  barriers | tmem base]
    unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_norm = reinterpret_cast<float*>(ring + (size_t)S * stage_bytes);  // [slot][0:64 |a|, 64:128 1/|a|]
    u64* bars = reinterpret_cast<u64*>(s_norm + kTcNormSlots * 2 * kTcN);
    u64* full_bar = bars;                 // [S]   TMA -> MMA
    u64* empty_bar = bars + S;            // [S]   MMA -> TMA
    u64* accf_bar = bars + 2 * S;         // [2]   MMA -> epilogue
    u64* acce_bar = bars + 2 * S + 2;     // [2]   epilogue -> MMA
    u64* aready_bar = bars + 2 * S + 4;   // [1]   queries are in TMEM
    u64* nfull_bar = bars + 2 * S + 5;                     // [kTcNormSlots] TMA -> epilogue (norm blocks)
    u64* nempty_bar = bars + 2 * S + 5 + kTcNormSlots;     // [kTcNormSlots] epilogue -> TMA
    u64* peerfull_bar = bars + 2 * S + 5 + 2 * kTcNormSlots;  // [S]   pair mode, leader: the peer's half landed
    u32* s_tmem = reinterpret_cast<u32*>(bars + 3 * S + 5 + 2 * kTcNormSlots);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (u32 i = 0; i < S; i++) {
            mbar_init(smem_u32(full_bar + i), 1);
            // multicast mode: one tcgen05.commit from every CTA of the cluster; pair mode: the leader's
            mbar_init(smem_u32(empty_bar + i), PAIR ? 1 : C);
            mbar_init(smem_u32(peerfull_bar + i), 1);
        }
        for (u32 i = 0; i < 2; i++) {
            mbar_init(smem_u32(accf_bar + i), 1);
            // pair: both CTAs' epilogues arrive at the leader (the peer's with one remote arrive per warp)
            mbar_init(smem_u32(acce_bar + i), PAIR ? kTcQ + 4 : kTcQ);
        }
        mbar_init(smem_u32(aready_bar), PAIR ? 2 * kTcQ : kTcQ);
        for (u32 i = 0; i < kTcNormSlots; i++) {
            mbar_init(smem_u32(nfull_bar + i), 1);
            mbar_init(smem_u32(nempty_bar + i), kTcQ);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {  // whole warp: allocate all 512 TMEM columns (1 CTA per SM)
        if (PAIR) {  // both CTAs of the pair issue the 2-CTA allocation, same warp, same destination
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (C > 1) cluster_sync_all();  // barrier inits visible to the peers before any remote arrive
    else __syncthreads();
    tc_fence_after();
    const u32 tmem = *s_tmem;
    const u32 tmem_a = tmem + kTcAccCols;  // A operand starts after the accumulators

    // tile sequence of this CTA: its tiles in order, the first `learn` of them bound-only, then those
    // same tiles once more at the end
    const u32 crank = C > 1 ? cluster_ctarank() : 0;
    const u32 cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    const u32 my_tiles = cluster_id < a.n_tiles ? (a.n_tiles - 1 - cluster_id) / n_clusters + 1 : 0;
    const u32 learn = min(kTcLearnTiles, my_tiles);
    const u32 n_seq = my_tiles + learn;

    if (warp == 4) {
        // ================= TMA producer =================
        if (lane == 0) {
            const u64 policy = make_evict_first_policy();
            u32 slot = 0, phase = 0;
            for (u32 sq = 0; sq < n_seq; sq++) {
                const u32 li = sq < my_tiles ? sq : sq - my_tiles;
                const size_t tile = (size_t)cluster_id + (size_t)li * n_clusters;
                {   // the tile's 64 norms and inverse norms (2 x 256 bytes)
                    const u32 ns = sq % kTcNormSlots;
                    mbar_wait(smem_u32(nempty_bar + ns), ((sq / kTcNormSlots) & 1) ^ 1);
                    mbar_expect_tx(smem_u32(nfull_bar + ns), 2 * kTcN * sizeof(float));
                    float* dst = s_norm + (size_t)ns * 2 * kTcN;
                    bulk_g2s(smem_u32(dst), a.norms + tile * a.norm_stride, kTcN * sizeof(float),
                             smem_u32(nfull_bar + ns), policy);
                    bulk_g2s(smem_u32(dst + kTcN), a.inv_norms + tile * a.norm_stride, kTcN * sizeof(float),
                             smem_u32(nfull_bar + ns), policy);
                }
                for (u32 kb0 = 0; kb0 < KB; kb0 += SKB) {  // one copy per stage (several column blocks)
                    const u32 bytes = min(SKB, KB - kb0) * kTileBytes;
                    // every CTA of the cluster has retired its MMAs on this slot
                    if (C > 1) mbar_wait_cluster(smem_u32(empty_bar + slot), phase ^ 1);
                    else mbar_wait(smem_u32(empty_bar + slot), phase ^ 1);
                    if (kb0 == 0) BSR_TL(sq, 0);
                    const unsigned char* src = a.tiles + (tile * KB + kb0) * kTileBytes;
                    const u32 dst = smem_u32(ring + (size_t)slot * stage_bytes);
                    if (PAIR) {
                        // my 32 rows of every 64-column block: 4 KiB at offset crank * 4 KiB of the block
                        const u32 nkb = min(SKB, KB - kb0);
                        mbar_expect_tx(smem_u32(full_bar + slot), nkb * blk_bytes);
                        for (u32 kk = 0; kk < nkb; kk++)
                            bulk_g2s(dst + kk * blk_bytes, src + (size_t)kk * kTileBytes + crank * blk_bytes, blk_bytes,
                                     smem_u32(full_bar + slot), policy);
                    } else if (C > 1) {
                        // multicast my 1/C of the stage to all CTAs of the cluster
                        mbar_expect_tx(smem_u32(full_bar + slot), bytes);
                        const u32 part = bytes / C;
                        bulk_g2s_multicast(dst + crank * part, src + (size_t)crank * part, part,
                                           smem_u32(full_bar + slot), (unsigned short)((1u << C) - 1), policy);
                    } else {
                        mbar_expect_tx(smem_u32(full_bar + slot), bytes);
                        bulk_g2s(dst, src, bytes, smem_u32(full_bar + slot), policy);
                    }
                    if (++slot == S) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer (pair mode: the leader CTA issues for both) =================
        if (lane == 0 && PAIR && crank != 0) {
            // non-leader CTA of a pair: tell the leader whenever my half of a stage has landed
            u32 slot = 0, phase = 0;
            for (u32 li = 0; li < n_seq; li++) {
                for (u32 kb0 = 0; kb0 < KB; kb0 += SKB) {
                    mbar_wait(smem_u32(full_bar + slot), phase);
                    mbar_arrive_cluster(smem_u32(peerfull_bar + slot), 0);
                    if (++slot == S) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (!(PAIR && crank != 0)) {
            // the whole warp runs this loop (see elect_one); every lane polls the barriers
            if (PAIR) mbar_wait_cluster(smem_u32(aready_bar), 0);
            else mbar_wait(smem_u32(aready_bar), 0);
            tc_fence_after();
            const unsigned short both = (unsigned short)((1u << C) - 1);
            u32 slot = 0, phase = 0;
            for (u32 li = 0; li < n_seq; li++) {
                const u32 buf = li & 1;
                // the epilogue(s) drained this accumulator
                if (PAIR) mbar_wait_cluster(smem_u32(acce_bar + buf), ((li >> 1) & 1) ^ 1);
                else mbar_wait(smem_u32(acce_bar + buf), ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) BSR_TL(li, 1);
                const u32 d_tmem = tmem + buf * kTcN;
                for (u32 kb0 = 0; kb0 < KB; kb0 += SKB) {
                    mbar_wait(smem_u32(full_bar + slot), phase);
                    if (PAIR) mbar_wait_cluster(smem_u32(peerfull_bar + slot), phase);  // ... and the peer's half
                    tc_fence_after();
                    if (lane == 0 && kb0 == 0) BSR_TL(li, 2);
                    const u32 nkb = min(SKB, KB - kb0);
                    const u32 stage_addr = smem_u32(ring + (size_t)slot * stage_bytes);
                    // One elected lane issues the whole stage.  The loop is fully unrolled with
                    // compile-time operand offsets (descriptor + kk * block + k4 * 32 bytes, A column
                    // + 8 per K16) so the issue stream is UTCHMMA after UTCHMMA: with run-time loop
                    // arithmetic between them (R2UR, descriptor rebuilds) the issuing lane needed
                    // ~45 cycles per MMA against the tensor pipe's 32 (timeline, B200).
                    if (elect_one()) {
                        const u64 desc0 = make_sw128_desc(stage_addr);
                        const u32 a0 = tmem_a + kb0 * 32;
#pragma unroll
                        for (u32 kk = 0; kk < kTcStageKB; kk++) {
                            if (kk < nkb) {
#pragma unroll
                                for (u32 k4 = 0; k4 < 4; k4++) {  // 4 x K16 per 64-column block
                                    const u32 acc = (kk | k4) != 0 ? 1u : (kb0 != 0 ? 1u : 0u);
                                    const u64 desc = desc0 + (u64)(kk * (kBlk >> 4) + k4 * 2);
                                    if (PAIR) tc_mma_ts_pair(d_tmem, a0 + (kk * 4 + k4) * 8, desc, kIdescBf16M256N64, acc);
                                    else tc_mma_ts(d_tmem, a0 + (kk * 4 + k4) * 8, desc, kIdescBf16M128N64, acc);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    // slot reusable once these MMAs retire -- in every CTA that shares the stage
                    if (elect_one()) {
                        if (PAIR) tc_commit_pair_multicast(smem_u32(empty_bar + slot), both);
                        else if (C > 1) tc_commit_multicast(smem_u32(empty_bar + slot), both);
                        else tc_commit(smem_u32(empty_bar + slot));
                    }
                    __syncwarp();
                    if (++slot == S) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
                // accumulator complete (pair mode: in both CTAs' tensor memory)
                if (elect_one()) {
                    if (PAIR) tc_commit_pair_multicast(smem_u32(accf_bar + buf), both);
                    else tc_commit(smem_u32(accf_bar + buf));
                }
                __syncwarp();
                if (lane == 0) BSR_TL(li, 3);
            }
        }
    } else {
        // ================= warps 0-3: query -> TMEM, then epilogue =================
        const u32 q = crank * kTcQ + tid;  // query index within the launch; TMEM lane == tid
        const u32 QS = C * kTcQ;           // queries per launch (stride of the per-query arrays)
        const bool active = q < a.nq_valid;
        const float* qp = a.queries + (size_t)(active ? q : 0) * a.dim;
        const u32 lane_addr = ((u32)(warp * 32) << 16);
        float qsq = 0.0f, qsq_exact = 0.0f;  // |b|^2 of the rounded query (bound scale) and of the query itself (distances)
        const bool vec_ok = (a.dim % 4 == 0) && ((reinterpret_cast<size_t>(qp) & 15) == 0);
        // Every lane reads its own query row (3 KiB at a 3 KiB stride), so this loop is bound by L2
        // round trips: four 32-element blocks are fetched (32 independent 128-bit loads) before the
        // first is converted -- tcgen05.st is an asm volatile the compiler will not move loads across.
        // bf16 rounding: cvt.rn.bf16x2.f32 (RNE; same values as f32_to_bf16_rne for every non-NaN input).
        const u32 n_cb = KB * 2;  // 32 elements -> 16 packed columns per store
        for (u32 cb0 = 0; cb0 < n_cb; cb0 += 4) {
            float f[4][32];
#pragma unroll
            for (u32 u = 0; u < 4; u++) {
                const u32 cb = cb0 + u;
                if (cb < n_cb) {
                    if (vec_ok && active && cb * 32 + 32 <= a.dim) {
                        const float4* q4 = reinterpret_cast<const float4*>(qp + cb * 32);
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const float4 t4 = q4[e];
                            f[u][4 * e] = t4.x; f[u][4 * e + 1] = t4.y; f[u][4 * e + 2] = t4.z; f[u][4 * e + 3] = t4.w;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; e++) {
                            const u32 c = cb * 32 + e;
                            f[u][e] = (active && c < a.dim) ? qp[c] : 0.0f;
                        }
                    }
                }
            }
#pragma unroll
            for (u32 u = 0; u < 4; u++) {
                const u32 cb = cb0 + u;
                if (cb < n_cb) {
                    u32 r[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[e]) : "f"(f[u][2 * e + 1]), "f"(f[u][2 * e]));
                        const float flo = bf16_lo(r[e]), fhi = bf16_hi(r[e]);
                        qsq = __fadd_rn(qsq, __fmul_rn(flo, flo));  // |b| over the rounded query, sequential
                        qsq = __fadd_rn(qsq, __fmul_rn(fhi, fhi));
                        // src/metrics.rs:155 on the query as given: columns in order, separately rounded
                        qsq_exact = __fadd_rn(qsq_exact, __fmul_rn(f[u][2 * e], f[u][2 * e]));
                        qsq_exact = __fadd_rn(qsq_exact, __fmul_rn(f[u][2 * e + 1], f[u][2 * e + 1]));
                    }
                    tmem_st16(tmem_a + lane_addr + cb * 16, r);
                }
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        if (PAIR && crank != 0) mbar_arrive_cluster(smem_u32(aready_bar), 0);  // the leader issues the MMAs
        else mbar_arrive(smem_u32(aready_bar));
        const float mag_b = __fsqrt_rn(qsq);
        const float mag_b_exact = __fsqrt_rn(qsq_exact);
        // The identity shortcut (src/metrics.rs:148-151) precedes every other rule.  Two vectors
        // within 1e-10 per element have sim ~ 1 unless the query itself is tiny (|b| < 1e-6): such
        // a query is never pruned and checks every row exactly.
        // (a NaN / inf element makes every score NaN: such a query prunes nothing either)
        const bool tiny_q = !(mag_b >= 1e-6f) || !(mag_b_exact >= 1e-6f) || !(mag_b_exact < INFINITY);

        // classes: (row position in its tile) x (CTA group); this CTA owns 64 of the kc classes
        const u32 T = (a.k + kTcN - 1) / kTcN, kc = T * kTcN;
        const u32 group = cluster_id % T;
        u32* g_gm = a.g_gm + q;  // class c of my query: g_gm[c * QS]  (a warp touches 128 contiguous bytes)
        float gmr[kTcN];  // my 64 class maxima, in registers
#pragma unroll
        for (u32 j = 0; j < kTcN; j++) gmr[j] = -INFINITY;

        // survivor slots of this (CTA, query): the upper half doubles as scratch for a compaction
        const u32 cap = a.surv_cap;
        u64* surv_buf = a.survivors + ((size_t)q * n_clusters + cluster_id) * (2 * (size_t)cap);
        u32 surv_cnt = 0;
        float thr = -INFINITY;    // prune rows with t < thr, t = S / |a|
        u32 next_refresh = max(learn >> 1, 1u);  // refresh the bound after this many tiles of the sequence

        for (u32 sq = 0; sq < n_seq; sq++) {
            const u32 buf = sq & 1;
            const u32 li = sq < my_tiles ? sq : sq - my_tiles;
            const bool learn_only = sq < learn;
            const u32 tile = cluster_id + li * n_clusters;
            const u32 row0 = tile * kTcN;
            const u32 ns = sq % kTcNormSlots;
            const float* na_s = s_norm + (size_t)ns * 2 * kTcN;
            const float4* inv4 = reinterpret_cast<const float4*>(na_s + kTcN);

            if (tid == 0) BSR_TL(sq, 4);
            mbar_wait(smem_u32(accf_bar + buf), (sq >> 1) & 1);
            tc_fence_after();
            if (tid == 0) BSR_TL(sq, 5);
            float v[kTcN];
            tmem_ld32(tmem + lane_addr + buf * kTcN, v);
            tmem_ld32(tmem + lane_addr + buf * kTcN + 32, v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            // scores are in registers: free the accumulator (pair mode: at the leader, which issues)
            if (PAIR && crank != 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(smem_u32(acce_bar + buf), 0);
            } else {
                mbar_arrive(smem_u32(acce_bar + buf));
            }
            mbar_wait(smem_u32(nfull_bar + ns), (sq / kTcNormSlots) & 1);
            if (tid == 0) BSR_TL(sq, 6);

            const u32 n_here = min(kTcN, a.n_rows - row0);
            // inactive lanes (no query) never survive anything
            const float pass_thr = (learn_only || !active) ? INFINITY : thr;
            if (n_here < kTcN) {
                // last tile of the store: padding rows score NaN -- FMNMX leaves the class maxima
                // alone, the NaN forces pass 1b, and the padding bits are dropped there
#pragma unroll
                for (u32 j = 0; j < kTcN; j++)
                    if (j >= n_here) v[j] = __uint_as_float(0x7FC00000u);
            }
            // pass 1 -- branch-free and unrolled, so the 64 independent chains overlap.  Per score: one
            // multiply, one FMNMX into its class maximum, and half a 3-input NaN-propagating FMNMX3
            // into the best score of its 16-row group.  Only where a group's best reaches the bound
            // (or is NaN) does the warp build that group's survivor bits (pass 1b).
            float gmax[4];
#pragma unroll
            for (u32 g = 0; g < 4; g++) {
                float m = -INFINITY;
#pragma unroll
                for (u32 j4 = g * 4; j4 < g * 4 + 4; j4++) {
                    const float4 iv = inv4[j4];
                    const float t0 = v[j4 * 4] * iv.x, t1 = v[j4 * 4 + 1] * iv.y;
                    const float t2 = v[j4 * 4 + 2] * iv.z, t3 = v[j4 * 4 + 3] * iv.w;
                    gmr[j4 * 4] = fmaxf(gmr[j4 * 4], t0);  // a NaN score leaves the class maximum alone
                    gmr[j4 * 4 + 1] = fmaxf(gmr[j4 * 4 + 1], t1);
                    gmr[j4 * 4 + 2] = fmaxf(gmr[j4 * 4 + 2], t2);
                    gmr[j4 * 4 + 3] = fmaxf(gmr[j4 * 4 + 3], t3);
                    m = fmax3_nan(m, t0, t1);
                    m = fmax3_nan(m, t2, t3);
                }
                gmax[g] = m;
            }
            const float tmax = fmax3_nan(fmax3_nan(gmax[0], gmax[1], gmax[2]), gmax[3], gmax[3]);
            // NaN scores always survive (their exact distance decides)
            if (!learn_only && __any_sync(0xFFFFFFFFu, !(tmax < pass_thr))) {
                // pass 2 -- only the 16-row groups where some lane of the warp can have a survivor:
                // re-test each score and append the survivors (compile-time register index, predicated
                // stores: no per-survivor loop, no register select)
                const u32 n_mine = active ? n_here : 0u;  // padding rows and idle lanes never append
#pragma unroll
                for (u32 g = 0; g < 4; g++) {
                    if (__any_sync(0xFFFFFFFFu, !(gmax[g] < pass_thr))) {
                        if (surv_cnt + 16 > cap) {
                            // overflow (adversarial row order only): keep the exact best k
                            compact_survivors(a, surv_buf, surv_cnt, qp, mag_b_exact, surv_buf + cap);
                            surv_cnt = a.k;
                        }
#pragma unroll
                        for (u32 j4 = g * 4; j4 < g * 4 + 4; j4++) {
                            const float4 iv = inv4[j4];
                            const float ivs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                            for (u32 e = 0; e < 4; e++) {
                                const u32 j = j4 * 4 + e;
                                const float t = v[j] * ivs[e];
                                if (!(t < pass_thr) && j < n_mine) surv_buf[surv_cnt++] = pack_survivor(row0 + j, t);
                            }
                        }
                    }
                }
            }
            mbar_arrive(smem_u32(nempty_bar + ns));  // norm block consumed

            if (sq + 1 == next_refresh) {
                // Merge my class maxima with everyone else's and refresh the bound.  The bound improves
                // like log(rows seen), so the refresh points are spaced geometrically (x 1.5).  Branch-free:
                // 64 coalesced L2 reads per class group, a predicated RED.MAX per improved class.
                // First points: learn/2 (publish early, so that nobody finds an unpublished class -- an
                // unpublished class means "no bound", i.e. every row survives -- when it reads at
                // `learn`), learn, learn + 1 (picks up what the others published at `learn`).
                next_refresh = next_refresh < learn ? learn
                                                    : max(next_refresh + 1, next_refresh + (next_refresh >> 1) - (next_refresh == learn ? (next_refresh >> 1) : 0u));
                if (active && !tiny_q) {
                    float bound = INFINITY;
                    for (u32 c0 = 0; c0 < kc; c0 += kTcN) {  // one 64-class group at a time
                        u32 their_bits[kTcN];
#pragma unroll
                        for (u32 i = 0; i < kTcN; i++) their_bits[i] = __ldcg(g_gm + (size_t)(c0 + i) * QS);  // independent, coalesced L2 reads
                        if (c0 == group * kTcN) {  // the group this CTA contributes to
#pragma unroll
                            for (u32 i = 0; i < kTcN; i++) {
                                // 0 = nobody has published this class yet
                                const float best = their_bits[i] ? f32_from_ordered(their_bits[i]) : -INFINITY;
                                red_max_if(g_gm + (size_t)(c0 + i) * QS, f32_ordered(gmr[i]), gmr[i] > best);
                                gmr[i] = fmaxf(gmr[i], best);
                                bound = fminf(bound, gmr[i]);
                            }
                        } else {
#pragma unroll
                            for (u32 i = 0; i < kTcN; i++)
                                bound = fminf(bound, their_bits[i] ? f32_from_ordered(their_bits[i]) : -INFINITY);
                        }
                    }
                    // kc >= k distinct rows reach `bound` in the FILTER's arithmetic; a row can only be dropped
                    // when even twice the worst-case filter error cannot lift it over them (a.margin, in units
                    // of |b| for cosine -- t = S / |a| -- and of |a|max |b| <= ... for BSR_DOT see the host side)
                    thr = (bound == -INFINITY) ? -INFINITY
                                               : bound - a.margin * mag_b_exact - 1e-6f * fabsf(bound);
                }
            }
            if (tid == 0) BSR_TL(sq, 7);
        }
        // exact keys + selection of the survivors happen in select_rescore_kernel (one CTA per query)
        if (active) {
            a.surv_count[(size_t)q * n_clusters + cluster_id] = surv_cnt;
            if (cluster_id == 0) a.q_mag[q] = mag_b_exact;
        }
    }

    tc_fence_before();
    if (C > 1) cluster_sync_all();  // no CTA may leave while a peer can still write into its shared memory
    else __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

#ifndef BSR_TEMPLATES_ONLY
// ascending bitonic sort of P (power of two) keys in shared memory by the whole CTA
__device__ __forceinline__ void cta_sort_keys(u64* buf, u32 P) {
    for (u32 size = 2; size <= P; size <<= 1) {
        for (u32 d = size >> 1; d >= 1; d >>= 1) {
            const u32 ld = 31 - __clz(d);
            for (u32 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
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
}

// Per query, after the scan (one CTA per query of the launch):
//  1. the k-th largest filter score t among ALL the query's survivors (every cluster's list) certifies
//     k distinct rows at least that good -- a far tighter bound than the running one the scan pruned
//     with (which sits near rank kc * H(kc)).  Radix selection over the order-preserving bit image of
//     t: 32 counting passes over the values held in shared memory (re-read from the lists when they do
//     not fit).
//  2. only survivors within the margin of that bound (~2 % of them; NaN scores always) can be in the
//     exact top-k: they are collected, get their exact key (the reference's distance over the stored
//     row), are sorted and folded into the query's running best K -- in chunks of kSelChunk, so that
//     any number of them (a query without a usable direction keeps every survivor) is handled.
//  3. the K best keys ARE the query's final candidate list: no per-cluster lists, no merge tree.
// Dynamic shared memory: [running K keys | union{values for step 1, chunk of step 2}].
constexpr u32 kThrSmemValues = 24576;  // 96 KiB: two CTAs per SM (top-100 on 10 M rows: ~9-12 k values per query)
constexpr u32 kSelChunk = 2048;        // >= BSR_MAX_TOP_K padded (1024): a chunk's best K merge with the running K
constexpr u32 kSelThreads = 512;
constexpr u32 kSelWarps = kSelThreads / 32;
__global__ void __launch_bounds__(kSelThreads) select_rescore_kernel(const TensorScanArgs a, u32 n_cta, u64* out_keys) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    const u32 K = a.k_pad;
    u64* running = reinterpret_cast<u64*>(sel_smem);              // [K]
    u64* chunk = running + K;                                      // [kSelChunk], contiguous with `running`
    u32* s_vals = reinterpret_cast<u32*>(running + K);             // [kThrSmemValues], dead before `chunk` is used
    __shared__ u32 s_count[kSelWarps];
    __shared__ u32 s_off[512];                                     // list offsets (n_cta <= 148 today)
    __shared__ u32 s_n;
    const u32 q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 cap = a.surv_cap;
    const u64* base = a.survivors + (size_t)q * n_cta * (2 * (size_t)cap);
    const u32* counts = a.surv_count + (size_t)q * n_cta;
    const float* qp = a.queries + (size_t)q * a.dim;
    const float mag_b = a.q_mag[q];

    for (u32 j = tid; j < K; j += kSelThreads) running[j] = kKeyMax;
    if (tid == 0) {  // exclusive prefix of the list lengths
        u32 acc = 0;
        for (u32 l = 0; l < n_cta; l++) {
            if (l < 512) s_off[l] = acc;
            acc += min(counts[l], cap);
        }
        s_n = acc;
    }
    __syncthreads();
    const u32 total = s_n;
    const bool in_smem = total <= kThrSmemValues && n_cta <= 512;
    // ---- step 1: the k-th largest score ---------------------------------------------------------
    if (in_smem) {  // one warp per list: no barrier while gathering
        for (u32 l = warp; l < n_cta; l += kSelThreads / 32) {
            const u32 n = min(counts[l], cap), off = s_off[l];
            for (u32 j = lane; j < n; j += 32) {
                const float t = __uint_as_float((u32)base[(size_t)l * (2 * (size_t)cap) + j]);
                s_vals[off + j] = (t == t) ? f32_ordered(t) : 0u;  // a NaN score certifies nothing: below every number
            }
        }
    }
    __syncthreads();
    auto count_ge = [&](u32 x) -> u32 {
        u32 c = 0;
        if (in_smem) {
            for (u32 j = tid; j < total; j += kSelThreads) c += s_vals[j] >= x ? 1u : 0u;
        } else {
            for (u32 l = 0; l < n_cta; l++) {
                const u32 n = min(counts[l], cap);
                for (u32 j = tid; j < n; j += kSelThreads) {
                    const float t = __uint_as_float((u32)base[(size_t)l * (2 * (size_t)cap) + j]);
                    c += ((t == t) ? f32_ordered(t) : 0u) >= x ? 1u : 0u;
                }
            }
        }
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        __syncthreads();  // s_count of the previous pass has been read by everyone
        if (lane == 0) s_count[warp] = c;
        __syncthreads();
        u32 sum = 0;
#pragma unroll
        for (u32 w = 0; w < kSelWarps; w++) sum += s_count[w];
        return sum;
    };
    float thr = -INFINITY;
    const u32 n_numbers = count_ge(1u);
    if (n_numbers >= a.k) {  // at least k survivors with a number for a score
        u32 ans = 0;
        for (int bit = 31; bit >= 0; bit--) {
            const u32 cand = ans | (1u << bit);
            if (count_ge(cand) >= a.k) ans = cand;
        }
        const float kth = f32_from_ordered(ans);  // k-th largest score
        thr = kth - a.margin * mag_b - 1e-6f * fabsf(kth);
        if (!(mag_b >= 1e-6f) || !(mag_b < INFINITY)) thr = -INFINITY;  // no usable direction: every survivor is rescored
    }
    // how many survivors pass (t >= thr, and every NaN score): usually a few hundred -> one chunk, collected
    // without a single barrier
    const u32 n_pass = (thr == -INFINITY ? n_numbers : count_ge(f32_ordered(thr))) + (total - n_numbers);
    __syncthreads();  // s_vals is dead from here on: `chunk` takes its place

    // ---- step 2: exact keys of the survivors that pass, folded into the running best K ------------
    auto process_chunk = [&](u32 n) {  // chunk[0..n) raw survivors -> running = K best of running U chunk
        u32 P = K;
        while (P < n) P <<= 1;  // <= kSelChunk
        for (u32 j = tid; j < P; j += kSelThreads) {
            u64 key = kKeyMax;
            if (j < n) {
                const u64 raw = chunk[j];
                key = survivor_key(a, raw, __ldg(a.norms + (u32)(raw >> 32)), qp, mag_b);
            }
            chunk[j] = key;  // entry j is read and written by the same thread
        }
        __syncthreads();
        cta_sort_keys(chunk, P);
        cta_merge_lists(running, 2, K);  // lists {running, chunk[0..K)} are contiguous; ends with __syncthreads()
    };
    if (tid == 0) s_n = 0;
    __syncthreads();
    if (n_pass <= kSelChunk) {
        for (u32 l = warp; l < n_cta; l += kSelWarps) {
            const u32 n = min(counts[l], cap);
            for (u32 j0 = 0; j0 < n; j0 += 32) {
                const u32 j = j0 + lane;
                u64 raw = 0;
                bool pass = false;
                if (j < n) {
                    raw = base[(size_t)l * (2 * (size_t)cap) + j];
                    const float t = __uint_as_float((u32)raw);
                    pass = !(t < thr);
                }
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
                u32 off = 0;
                if (lane == 0 && bal) off = atomicAdd(&s_n, (u32)__popc(bal));
                off = __shfl_sync(0xFFFFFFFFu, off, 0);
                if (pass) chunk[off + __popc(bal & ((1u << lane) - 1u))] = raw;
            }
        }
    } else
    for (u32 l0 = 0; l0 < n_cta; l0 += kSelWarps) {  // a round of lists at a time, one warp each
        const u32 l = l0 + warp;
        const u32 n = l < n_cta ? min(counts[l], cap) : 0u;
        u32 n_max = n;  // the longest of the eight lists bounds the number of rounds (warp-uniform per warp, CTA-uniform below)
        for (u32 j0 = 0;; j0 += 32) {
            // CTA-uniform loop: every warp takes part in the barriers until all eight lists are exhausted
            const u32 j = j0 + lane;
            u64 raw = 0;
            bool pass = false;
            if (j < n) {
                raw = base[(size_t)l * (2 * (size_t)cap) + j];
                const float t = __uint_as_float((u32)raw);
                pass = !(t < thr);
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
            u32 off = 0;
            if (lane == 0 && bal) off = atomicAdd(&s_n, (u32)__popc(bal));
            off = __shfl_sync(0xFFFFFFFFu, off, 0);
            if (pass) chunk[off + __popc(bal & ((1u << lane) - 1u))] = raw;  // room: at most 256 appends per round, see below
            const bool more = __syncthreads_or(j0 + 32 < n_max);
            // a full round of warps adds at most kSelThreads entries: process before the next one could overflow
            if (s_n > kSelChunk - kSelThreads) {
                const u32 cnt = s_n;
                __syncthreads();
                process_chunk(cnt);
                if (tid == 0) s_n = 0;
                __syncthreads();
            }
            if (!more) break;
        }
    }
    __syncthreads();
    if (s_n) {
        const u32 cnt = s_n;
        __syncthreads();
        process_chunk(cnt);
    }
    // ---- step 3: the query's final list ----------------------------------------------------------
    u64* out = out_keys + (size_t)q * K;
    for (u32 j = tid; j < K; j += kSelThreads) out[j] = j < a.k ? running[j] : kKeyMax;  // slots past top_k are published empty
}

// largest finite |a| (norms are >= 0: their bit patterns order like the values)
__global__ void max_norm_kernel(const float* __restrict__ norms, u32 n, u32* __restrict__ out_bits) {
    float m = 0.0f;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = norms[i];
        if (v < INFINITY) m = fmaxf(m, v);
    }
    for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

// fp32 store -> its bf16 shadow in the tensor path's tile layout (rows [row_lo, row_hi), whole 16-byte
// chunks: thread t owns destination chunk t of the tile range), plus 1/|a| for the filter's score scale.
__global__ void shadow_from_panels_kernel(const float4* __restrict__ panels, uint4* __restrict__ tiles,
                                          const float* __restrict__ norms, float* __restrict__ inv_norms,
                                          u32 row_lo, u32 row_hi, u32 dim, u32 dq, u32 KB) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 row, col;
    bf16_chunk_coords(t, row_lo >> 6, KB, row, col);
    if (row >= ((row_hi + 63u) & ~63u)) return;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = 0.0f;
    // every row of the first tile is (re)converted, also those below row_lo: same values as before
    if (row < row_hi) {
        const float4* p = panels + (size_t)(row >> 5) * dq * 32 + (row & 31u);
        if (col < dim) {  // columns past dim are zero in the panels as well
            const float4 x = p[(size_t)(col >> 2) * 32];
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
            if (col + 4 < dq * 4) {
                const float4 y = p[(size_t)((col >> 2) + 1) * 32];
                v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
            }
        }
        if (col == 0) {
            const float na = norms[row];
            inv_norms[row] = (na > 0.0f && na < INFINITY) ? __fdiv_rn(1.0f, na) : 0.0f;
        }
    }
    tiles[(size_t)(row_lo >> 6) * KB * 512 + t] = pack8_bf16(v);
}

#endif  // BSR_TEMPLATES_ONLY

}  // namespace bsr
