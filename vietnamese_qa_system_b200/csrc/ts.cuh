// ts.cuh -- K2c + K3: the large-batch (GEMM-regime) tcgen05 kernel: queries resident in TENSOR
// MEMORY as the MMA's A operand, documents streamed through the whole of shared memory.
//
// Replaces faiss IndexFlatIP search under txtai ann.search (heavy_ranker.py:98,100) for query
// batches beyond what mma.cuh can keep resident in shared memory.  There the queries (hi + lo,
// 3 KB each at dim 768) compete with the TMA ring for the SM's 227 KB, and the ring's size over the
// memory latency bounds how fast one SM can ingest documents (Little's law: measured 1.5x the HBM
// share, whatever feeds it -- L2 re-reads or cluster multicast).  Here:
//   * A = the query block, M = 128 query rows x K = dim, written once into TMEM with tcgen05.st
//     (dim/2 of the 512 columns: 384 at dim 768) and read by `tcgen05.mma [d], [a_tmem], b_desc`;
//   * B = 64-document x 64-column boxes (8 KB, 128B-swizzled, TMA) -- shared memory holds nothing
//     but the ring (~200 KB in flight per SM) and the per-thread lists;
//   * D = 128 queries x 64 documents fp32 in the remaining TMEM columns, (512 - dim/2)/64
//     accumulator stages (2 at dim 768), so the epilogue of tile t overlaps the MMAs of tile t+1;
//   * epilogue: thread = one QUERY row: tcgen05.ld gives it its 64 document scores, it keeps a
//     private threshold in a register and a private sorted top-k list in shared memory -- no
//     cross-thread traffic at all; thresholds are shared GPU-wide exactly as in mma.cuh.
// With `split` the 128 rows are 64 queries as hi rows (0..63) and lo rows (64..127); the lo rows'
// scores reach the hi rows through a 16 KB shared-memory exchange per tile.
// One HBM pass serves 128 (or 64) queries per CTA, x cluster size with TMA multicast.
//
// Roofline: HBM up to ~128 queries per pass (bytes = n_rows*dim*2 per launch), tensor pipe beyond.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "mma.cuh"
#include "ptx.cuh"

namespace vqa {

constexpr int kTsRows = 128;       // query rows per CTA (UMMA M)
constexpr int kTsDocs = 64;        // documents per MMA tile (UMMA N)
constexpr int kTsBoxBytes = kTsDocs * kBlockK * 2;  // 8 KB per TMA box

struct TsParams {
    const float *q;
    long long q_stride;
    int nq;        // queries in this launch
    int per_cta;   // queries per CTA: 128, or 64 with split
    int split;
    int a_fp16;    // must be 0: mixing an fp16 A with bf16 B in one kind::f16 MMA is an illegal instruction on sm_100a
    int k;
    long long n_rows;
    int dim;       // multiple of 64, <= 768
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
    int n_tiles;   // tiles of 64 documents
    int n_stages;
    int kps;
    unsigned long long tma_policy;
    unsigned long long *tau_g;
    uint32_t epoch;
    int n_groups;
    int multicast;
    int ks;        // QS variants: the LAST ks 64-column blocks of the query block live in shared memory, not TMEM
    // Diagnostic (vqa_debug_timeline), normally nullptr: per CTA 32 uint64 -- 0 entry / 15 exit (%globaltimer ns),
    // SM-clock cycles: 1 producer waiting for a free ring stage, 2 MMA warp waiting for documents, 3 MMA warp waiting
    // for a free accumulator stage, 4 MMA warp's whole loop, 5 epilogue warp 0 waiting for an accumulator, 6 its whole
    // loop, 7 its list flushes (count), 8 tiles of this CTA, 9 producer's whole loop
    unsigned long long *timeline;
};

constexpr int kTsQBlockBytes = kTsRows * kBlockK * 2;  // 16 KB: 128 query rows x one 64-column block, 128B-swizzled

// register-list variants: 16 (k <= 16) or 32 (k <= 32) entries per query row; 0 = list in shared memory
inline int ts_reg_list_len(int k) { return k <= 16 ? 16 : (k <= 32 ? 32 : 0); }

// rows that own a shared-memory list / candidate buffer: 64 when at most 64 rows of the CTA are queries (hi/lo
// rows, or -- QS variants only -- a launch of at most 64 queries) and the lists are in shared memory (k > 32)
inline int ts_list_rows(int k, int split, int nq, int qs) {
    return (ts_reg_list_len(k) == 0 && (split || (qs && nq <= 64))) ? 64 : kTsRows;
}

// [align slack][QS: ks query blocks x 16 KB][ring: boxes x 8 KB][barriers 1 KB][exchange 2 x 16 KB if split]
// [lists: rows x k x 8 B, or the 32-deep candidate buffers (rows x 32 x 8 B) of the register-list variants]
// shared-memory lists (k > 32) are kept for the live rows only: 64 with hi/lo rows, else 128; every
// variant also has 32-deep candidate buffers for those rows
// m64: the M = 64 variant (<= 64 queries per CTA): a shared-memory query block is 64 rows = 8 KB
inline size_t ts_smem_bytes_rt(int k, int boxes, int split, int ks = 0, int nq = 1 << 30, int qs = 0, int m64 = 0) {
    const int depth = ts_reg_list_len(k) > 0 ? 32 : k + 32;
    const int rows = ts_list_rows(k, split, nq, qs);
    return 1024 + (size_t)ks * (m64 ? kTsQBlockBytes / 2 : kTsQBlockBytes) + (size_t)boxes * kTsBoxBytes + 1024 +
           (split ? 2 * 64 * kTsDocs * 4 : 0) + (size_t)rows * depth * 8;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// QS ("query block partly in Shared memory", opt-in until timed on a B200): the last p.ks 64-column blocks of
// the query block are staged in shared memory (K-major, 128B-swizzled, 16 KB each) and multiplied with the
// shared-memory-A form of the MMA into the same accumulator; only the first dim/64 - ks blocks occupy TMEM.
// That (a) fits dim 1024 (BASELINE configs[3]: 12 blocks = 384 columns in TMEM + 4 blocks = 64 KB in shared
// memory, 2 accumulator stages) and (b) is a knob at dim 768: ks = 4 leaves 256 TMEM columns = 4 accumulator
// stages instead of 2.  (Default since round 2; QS = false is round 1's kernel.)
// M64 (QS variants, <= 64 queries per CTA, no hi/lo rows): M = 64 instructions.  The 64 query rows sit in lanes 0..15
// of each of the four 32-lane quarters of tensor memory (row r <-> lane 32 * (r / 16) + r % 16, pinned on the hardware
// by tools/m64_probe.cu), so each instruction reads half the A bytes from tensor memory -- the read port (64 B/cycle)
// is what paces the M = 128 x N = 64 tiles -- and a shared-memory query block is 8 KB instead of 16.
template <bool DOC_BF16, int KL, bool QS = false, bool M64 = false>
__global__ void __launch_bounds__(kMmaThreads, 1)
ts_topk_kernel(const __grid_constant__ CUtensorMap tmap_docs, const TsParams p) {
    static_assert(!M64 || QS, "M = 64 is a QS variant");
    constexpr int QBLK = M64 ? kTsQBlockBytes / 2 : kTsQBlockBytes;   // bytes of one shared-memory query block
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int KB = p.dim / kBlockK;
    const int S = p.n_stages;
    const int KPS = p.kps;
    const int KG = KB / KPS;
    const uint32_t stage_bytes = (uint32_t)KPS * kTsBoxBytes;
    const int KSB = QS ? p.ks : 0;                  // query blocks in shared memory
    const int KT = KB - KSB;                        // query blocks in tensor memory
    const int ACOLS = QS ? KT * (kBlockK / 2) : p.dim / 2;  // TMEM columns of the query block
    const int AS = (QS && (512 - ACOLS) / kTsDocs > kMaxAccStages) ? kMaxAccStages : (512 - ACOLS) / kTsDocs;  // accumulator stages
    unsigned char *q_smem = smem;                   // QS: KSB blocks of 128 query rows x 128 bytes
    unsigned char *a_smem = smem + (size_t)KSB * QBLK;  // document ring
    uint64_t *bars = reinterpret_cast<uint64_t *>(a_smem + (size_t)S * stage_bytes);
    uint64_t *full = bars;
    uint64_t *empty = bars + kMaxStages;
    uint64_t *tfull = bars + 2 * kMaxStages;
    uint64_t *tempty = tfull + kMaxAccStages;
    uint64_t *qready = tempty + kMaxAccStages;   // the query block is in place (one arrival per epilogue warp)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(qready + 1);
    float *xchg = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 1024);  // [2][64 docs][64 rows]
    unsigned char *lists = reinterpret_cast<unsigned char *>(xchg) + (p.split ? 2 * 64 * kTsDocs * 4 : 0);
    // per-thread sorted lists (KL == 0) and candidate buffers, entry-major so that a warp's accesses are
    // conflict free; LR = rows that own a list (64 when the upper 64 rows are the queries' lo parts)
    const int LR = (KL == 0 && (p.split || (QS && p.nq <= 64))) ? 64 : kTsRows;
    float *lst_s = reinterpret_cast<float *>(lists);                       // [k][LR]        (KL == 0)
    uint32_t *lst_i = reinterpret_cast<uint32_t *>(lst_s + (size_t)p.k * LR);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);
    const int grp = blockIdx.x % p.n_groups;
    const int stream0 = blockIdx.x / p.n_groups;
    const int n_streams = gridDim.x / p.n_groups;
    const int q0 = grp * p.per_cta;
    const int nq = p.nq - q0 < p.per_cta ? (p.nq - q0 > 0 ? p.nq - q0 : 0) : p.per_cta;
    const uint16_t cta_mask = (uint16_t)((1u << p.n_groups) - 1u);
    const int slice_rows = kTsDocs / (p.multicast ? p.n_groups : 1);
    const uint32_t idesc = (1u << 4) | ((p.a_fp16 ? 0u : (DOC_BF16 ? 1u : 0u)) << 7) | ((DOC_BF16 ? 1u : 0u) << 10) |
                           ((uint32_t)(kTsDocs >> 3) << 17) | ((uint32_t)((M64 ? 64 : kTsRows) >> 4) << 24);

    if (p.timeline != nullptr && tid == 0) p.timeline[(size_t)blockIdx.x * 32] = ptx::globaltimer_ns();
    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&tmap_docs);
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(full + s, 1);
            ptx::mbar_init(empty + s, p.multicast ? p.n_groups : 1);
        }
        for (int a = 0; a < AS; ++a) {
            ptx::mbar_init(tfull + a, 1);
            ptx::mbar_init(tempty + a, 4);
        }
        ptx::mbar_init(qready, 4);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    if (p.multicast) ptx::cluster_sync_all();
    if (warp == 4) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
        ptx::tc_fence_before_sync();
    }
    grid_launch_dependents();
    __syncthreads();  // TMEM base visible to everyone (the producer has not started yet: cheap, once)
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(kFullMask, *reinterpret_cast<volatile uint32_t *>(tmem_slot), 0);

    if (warp == 4) {
        // ===== TMA producer =====
        // (does not wait for the query block: documents start streaming now)
        uint32_t it = 0;
        const bool tl = p.timeline != nullptr;
        unsigned long long w_empty = 0, t_loop = tl ? ptx::sm_clock() : 0;
        for (int tile = stream0; tile < p.n_tiles; tile += n_streams) {
            for (int kg = 0; kg < KG; ++kg, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                const unsigned long long t_w = tl ? ptx::sm_clock() : 0;
                ptx::mbar_wait(empty + s, ph ^ 1);
                if (tl) w_empty += ptx::sm_clock() - t_w;
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(full + s, stage_bytes);
                    for (int j = 0; j < KPS; ++j) {
                        unsigned char *dst = a_smem + (size_t)s * stage_bytes + (size_t)j * kTsBoxBytes;
                        if (p.multicast)
                            ptx::tma_load_2d_multicast(dst + (size_t)grp * slice_rows * 128, &tmap_docs,
                                                       (kg * KPS + j) * kBlockK, tile * kTsDocs + grp * slice_rows,
                                                       full + s, cta_mask, p.tma_policy);
                        else
                            ptx::tma_load_2d(dst, &tmap_docs, (kg * KPS + j) * kBlockK, tile * kTsDocs, full + s,
                                             p.tma_policy);
                    }
                }
                __syncwarp();
            }
        }
        if (tl && lane == 0) {
            p.timeline[(size_t)blockIdx.x * 32 + 1] = w_empty;
            p.timeline[(size_t)blockIdx.x * 32 + 9] = ptx::sm_clock() - t_loop;
        }
    } else if (warp == 5) {
        // ===== MMA issuer: waits for the query block (mbarrier `qready`: the four epilogue warps arrive once their rows
        // are in tensor / shared memory), then D = Q * docs^T.  (Until round 2 this was a named bar.sync reached from two
        // code locations -- legal, but compute-sanitizer's synccheck reports it as block-level divergence.)
        ptx::mbar_wait(qready, 0);
        ptx::tc_fence_after_sync();
        uint32_t it = 0, lt = 0;
        const uint32_t ring = ptx::smem_u32(a_smem);
        const uint32_t qs_base = ptx::smem_u32(q_smem);
        const bool tl = p.timeline != nullptr;
        unsigned long long w_full = 0, w_tempty = 0, t_loop = tl ? ptx::sm_clock() : 0;
        for (int tile = stream0; tile < p.n_tiles; tile += n_streams, ++lt) {
            const int as = lt % AS;
            const uint32_t aph = (lt / AS) & 1;
            unsigned long long t_w = tl ? ptx::sm_clock() : 0;
            ptx::mbar_wait(tempty + as, aph ^ 1);
            if (tl) w_tempty += ptx::sm_clock() - t_w;
            ptx::tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)(ACOLS + as * kTsDocs);
            for (int kg = 0; kg < KG; ++kg, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                t_w = tl ? ptx::sm_clock() : 0;
                ptx::mbar_wait(full + s, ph);
                if (tl) w_full += ptx::sm_clock() - t_w;
                ptx::tc_fence_after_sync();
                if (ptx::elect_one()) {
                    for (int j = 0; j < KPS; ++j) {
                        const int kb = kg * KPS + j;
                        const uint64_t db0 = ptx::umma_desc_k_sw128(ring + (uint32_t)s * stage_bytes + (uint32_t)j * kTsBoxBytes);
                        if (QS && kb >= KT) {  // this block of the queries is in shared memory
                            const uint64_t da0 = ptx::umma_desc_k_sw128(qs_base + (uint32_t)(kb - KT) * QBLK);
#pragma unroll
                            for (int k4 = 0; k4 < kBlockK / 16; ++k4)  // +32 bytes along K = +2 in the address field
                                ptx::umma_f16(d_tmem, da0 + (uint64_t)(k4 * 2), db0 + (uint64_t)(k4 * 2), idesc,
                                              (kb | k4) != 0 ? 1u : 0u);
                            continue;
                        }
#pragma unroll
                        for (int k4 = 0; k4 < kBlockK / 16; ++k4)  // 16 K-elements = 8 TMEM columns of A
                            umma_f16_ts(d_tmem, tmem_base + (uint32_t)(kb * 32 + k4 * 8), db0 + (uint64_t)(k4 * 2), idesc,
                                        (kb | k4) != 0 ? 1u : 0u);
                    }
                    if (p.multicast) ptx::umma_commit_multicast(empty + s, cta_mask);
                    else ptx::umma_commit(empty + s);
                    if (kg == KG - 1) ptx::umma_commit(tfull + as);
                }
                __syncwarp();
            }
        }
        if (tl && lane == 0) {
            unsigned long long *o = p.timeline + (size_t)blockIdx.x * 32;
            o[2] = w_full;
            o[3] = w_tempty;
            o[4] = ptx::sm_clock() - t_loop;
            o[8] = lt;
        }
    } else {
        // ===== warps 0-3: thread = query row =====
        const int row = warp * 32 + lane;                       // TMEM lane
        const int qrow = M64 ? warp * 16 + (lane & 15) : (p.split ? (row & 63) : row);   // query of this row within the chunk
        const bool is_lo = !M64 && p.split && row >= 64;
        const bool live = qrow < nq && (!M64 || lane < 16);     // (M = 64: lanes 16..31 of every quarter are not rows)
        // 1. write this row of the query block into TMEM (16-bit pairs, K ascending)
        {
            const float *src = p.q + (long long)(q0 + (live ? qrow : 0)) * p.q_stride;
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
            // QS variants: the loads of four 16-column groups (32 float4) are issued together -- the classic loop
            // below pays one L2 round trip per group (24 at dim 768, ~20 us before the first MMA can be issued)
            constexpr int QG = QS ? 4 : 1;
            for (int c0 = 0; QS && c0 < ACOLS; c0 += 16 * QG) {
                float4 v[QG][8];
#pragma unroll
                for (int g = 0; g < QG; ++g)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        v[g][jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live && c0 + g * 16 < ACOLS)
                            v[g][jj] = *reinterpret_cast<const float4 *>(src + (c0 + g * 16 + jj * 2) * 2);
                    }
#pragma unroll
                for (int g = 0; g < QG; ++g) {
                    if (c0 + g * 16 >= ACOLS) break;  // warp-uniform
                    uint32_t w[16];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        float x[4] = {v[g][jj].x, v[g][jj].y, v[g][jj].z, v[g][jj].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float hi;
                            if (DOC_BF16 && !p.a_fp16) hi = __bfloat162float(__float2bfloat16_rn(x[e]));
                            else hi = __half2float(__float2half_rn(x[e]));
                            x[e] = is_lo ? x[e] - hi : hi;
                        }
                        if (DOC_BF16 && !p.a_fp16) {
                            w[2 * jj] = pack_bf16x2(x[0], x[1]);
                            w[2 * jj + 1] = pack_bf16x2(x[2], x[3]);
                        } else {
                            w[2 * jj] = pack_f16x2(x[0], x[1]);
                            w[2 * jj + 1] = pack_f16x2(x[2], x[3]);
                        }
                    }
                    tmem_st16(trow + (uint32_t)(c0 + g * 16), w);
                }
            }
            for (int c0 = 0; !QS && c0 < ACOLS; c0 += 16) {
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live) v = *reinterpret_cast<const float4 *>(src + (c0 + j) * 2);
                    float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float hi;
                        if (DOC_BF16 && !p.a_fp16) hi = __bfloat162float(__float2bfloat16_rn(x[e]));
                        else hi = __half2float(__float2half_rn(x[e]));
                        x[e] = is_lo ? x[e] - hi : hi;
                    }
                    if (DOC_BF16 && !p.a_fp16) {
                        w[j] = pack_bf16x2(x[0], x[1]);
                        w[j + 1] = pack_bf16x2(x[2], x[3]);
                    } else {
                        w[j] = pack_f16x2(x[0], x[1]);
                        w[j + 1] = pack_f16x2(x[2], x[3]);
                    }
                }
                tmem_st16(trow + (uint32_t)c0, w);
            }
            tmem_st_wait();
            if constexpr (QS) {
                // the remaining blocks of this row -> shared memory, K-major, 128-byte swizzle (16-byte chunk c of
                // row r at r * 128 + ((c ^ (r & 7)) << 4)), same 16-bit values as the TMEM part
                // (M = 64: lanes 16..31 own no row; they walk the loop with the stores switched off, so that the warp
                //  reaches the block-wide barrier below converged -- a divergent loop exit here made synccheck report
                //  the bar.sync as executed by a split warp)
                const bool owns_row = !M64 || lane < 16;
                for (int kb = KT; kb < KB; ++kb) {
                    unsigned char *tile = q_smem + (size_t)(kb - KT) * QBLK;
                    const int row = M64 ? qrow : warp * 32 + lane;   // row of the K-major shared-memory tile
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                        if (live) {
                            v0 = *reinterpret_cast<const float4 *>(src + kb * kBlockK + c * 8);
                            v1 = *reinterpret_cast<const float4 *>(src + kb * kBlockK + c * 8 + 4);
                        }
                        float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float hi;
                            if (DOC_BF16 && !p.a_fp16) hi = __bfloat162float(__float2bfloat16_rn(x[e]));
                            else hi = __half2float(__float2half_rn(x[e]));
                            x[e] = is_lo ? x[e] - hi : hi;
                        }
                        uint4 w;
                        if (DOC_BF16 && !p.a_fp16) w = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                        else w = make_uint4(pack_f16x2(x[0], x[1]), pack_f16x2(x[2], x[3]), pack_f16x2(x[4], x[5]), pack_f16x2(x[6], x[7]));
                        if (owns_row) *reinterpret_cast<uint4 *>(tile + row * 128 + ((c ^ (row & 7)) << 4)) = w;
                    }
                }
                ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(qready);   // -> the MMA warp; the epilogue warps do not wait for each other
        }
        // 2. private list + threshold
        float tau = (live && !is_lo) ? neg_inf() : __int_as_float(0x7f800000);
        unsigned long long *tg = (p.tau_g != nullptr && live && !is_lo) ? p.tau_g + q0 + qrow : nullptr;
        // KL > 0: the sorted list lives in REGISTERS (rs/ri); candidates that beat the threshold are
        // first appended to a private shared-memory buffer (two predicated stores, no divergence) and the
        // whole warp folds its buffers into the lists in lockstep when any lane's buffer is half full.
        // KL == 0 (k > 32): same buffers; the list is a worst-at-root binary heap in shared memory.
        constexpr int KLR = KL > 0 ? KL : 1;
        constexpr int CAP = 32;
        float rs[KLR];
        uint32_t ri[KLR];
        int cnt = 0;
        float worst_s = neg_inf();                  // KL == 0: the heap's root = the worst kept entry
        uint32_t worst_i = invalid_id<uint32_t>();
        const int lrow = M64 ? qrow : (row & (LR - 1));   // this thread's list / buffer column
        float *buf_s = KL > 0 ? lst_s : reinterpret_cast<float *>(lst_i + (size_t)p.k * LR);   // [CAP][LR]
        uint32_t *buf_i = reinterpret_cast<uint32_t *>(buf_s + CAP * LR);
        if constexpr (KL > 0) {
#pragma unroll
            for (int e = 0; e < KLR; ++e) {
                rs[e] = neg_inf();
                ri[e] = invalid_id<uint32_t>();
            }
        } else if (M64 ? lane < 16 : (!is_lo && (!QS || row < LR))) {  // (QS with <= 64 queries: rows 64.. own no list column;
                                                                         //  M = 64: the rows are lanes 0..15 of every quarter)
            for (int e = 0; e < p.k; ++e) {
                lst_s[e * LR + lrow] = neg_inf();
                lst_i[e * LR + lrow] = invalid_id<uint32_t>();
            }
        }
        auto flush = [&]() __attribute__((always_inline)) {
            if constexpr (KL > 0) {
                const int wmax = __reduce_max_sync(kFullMask, cnt);
                for (int b = 0; b < wmax; ++b) {
                    float cs = neg_inf();
                    uint32_t ci = invalid_id<uint32_t>();
                    if (b < cnt) {
                        cs = buf_s[b * LR + lrow];
                        ci = buf_i[b * LR + lrow];
                    }
                    bool bef[KLR];
#pragma unroll
                    for (int e = 0; e < KLR; ++e) bef[e] = ranks_before<uint32_t>(cs, ci, rs[e], ri[e]);
#pragma unroll
                    for (int e = KLR - 1; e >= 0; --e) {
                        const bool up = e > 0 ? bef[e - 1] : false;  // candidate lands above e: entry e-1 moves down
                        rs[e] = up ? rs[e > 0 ? e - 1 : 0] : (bef[e] ? cs : rs[e]);
                        ri[e] = up ? ri[e > 0 ? e - 1 : 0] : (bef[e] ? ci : ri[e]);
                    }
                }
                cnt = 0;
                // threshold = score of rank k-1 (p.k <= KL), once that slot is filled.  Written as reductions over
                // the whole (sorted) list: `if (e == p.k - 1) ts = rs[e]` is turned into rs[p.k - 1] by the compiler,
                // and ONE dynamic index sends both register arrays to local memory (128-byte stack frame, ~12 LDL +
                // 12 STL per tile and warp -- r2_ts_b256.ncu-rep: 14.8 M local requests per launch)
                float ts = __int_as_float(0x7f800000);
                int filled = 0;
#pragma unroll
                for (int e = 0; e < KLR; ++e) {
                    ts = fminf(ts, e < p.k ? rs[e] : __int_as_float(0x7f800000));
                    filled += (e < p.k && ri[e] != invalid_id<uint32_t>()) ? 1 : 0;
                }
                if (filled == p.k && ts > tau) {
                    tau = ts;
                    if (tg != nullptr) atomicMax(tg, tau_encode(ts, p.epoch));
                }
            } else {
                // The list is a binary HEAP in shared memory with the WORST kept entry at the root (parent
                // ranks after both children): a candidate that ranks before the root replaces it and sifts
                // down -- O(log k) dependent steps, usually one or two because a candidate that barely beats
                // the worst entry rarely beats that entry's children.  (Sorted insertion costs a ~k/2-entry
                // shift; the reduce kernel re-ranks the published candidates anyway.)
                const int wmax = __reduce_max_sync(kFullMask, cnt);
                const int k = p.k;
                for (int b = 0; b < wmax; ++b) {
                    if (b < cnt) {
                        const float cs = buf_s[b * LR + lrow];
                        const uint32_t ci = buf_i[b * LR + lrow];
                        if (ranks_before<uint32_t>(cs, ci, worst_s, worst_i)) {
                            int i = 0;
                            while (true) {
                                int c = 2 * i + 1;
                                if (c >= k) break;
                                float s1 = lst_s[c * LR + lrow];
                                uint32_t i1 = lst_i[c * LR + lrow];
                                if (c + 1 < k) {
                                    const float s2 = lst_s[(c + 1) * LR + lrow];
                                    const uint32_t i2 = lst_i[(c + 1) * LR + lrow];
                                    if (ranks_before<uint32_t>(s1, i1, s2, i2)) {  // child c+1 is the worse one
                                        c = c + 1;
                                        s1 = s2;
                                        i1 = i2;
                                    }
                                }
                                if (!ranks_before<uint32_t>(cs, ci, s1, i1)) break;  // candidate is no better: it stays here
                                lst_s[i * LR + lrow] = s1;                            // the worse child moves up
                                lst_i[i * LR + lrow] = i1;
                                i = c;
                            }
                            lst_s[i * LR + lrow] = cs;
                            lst_i[i * LR + lrow] = ci;
                            worst_s = lst_s[lrow];
                            worst_i = lst_i[lrow];
                        }
                    }
                }
                // threshold = score of the worst kept entry once all k slots are filled
                if (cnt > 0 && worst_i != invalid_id<uint32_t>() && worst_s > tau) {
                    tau = worst_s;
                    if (tg != nullptr) atomicMax(tg, tau_encode(worst_s, p.epoch));
                }
                cnt = 0;
            }
        };

        uint32_t lt = 0;
        const bool tl = p.timeline != nullptr && warp == 0;
        unsigned long long w_tfull = 0, n_flush = 0, t_loop = tl ? ptx::sm_clock() : 0;
        for (int tile = stream0; tile < p.n_tiles; tile += n_streams, ++lt) {
            const int as = lt % AS;
            const uint32_t aph = (lt / AS) & 1;
            unsigned long long graw = 0;
            if (tg != nullptr) graw = ld_volatile_u64(tg);
            const unsigned long long t_w = tl ? ptx::sm_clock() : 0;
            ptx::mbar_wait(tfull + as, aph);
            if (tl) w_tfull += ptx::sm_clock() - t_w;
            ptx::tc_fence_after_sync();
            if (tg != nullptr) tau = fmaxf(tau, tau_decode(graw, p.epoch));
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ACOLS + as * kTsDocs);
            const long long doc0 = (long long)tile * kTsDocs;
            const int ndoc = p.n_rows - doc0 < kTsDocs ? (int)(p.n_rows - doc0) : kTsDocs;
            float *xb = xchg + (size_t)(lt & 1) * 64 * kTsDocs;
            if constexpr (QS) {
                // QS variants: all 64 scores of the row -> registers, then the accumulator stage goes back to the
                // MMA warp BEFORE the scores are looked at (the classic loop below holds it through every flush).
                // Hot path: one max tree + one ballot per tile; the append / flush code exists once, in a rolled loop.
                float v[kTsDocs];
                {
                    uint32_t acc[4][16];
#pragma unroll
                    for (int c = 0; c < 4; ++c) ptx::tmem_ld16(taddr + c * 16, acc[c]);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[c * 16 + j] = __uint_as_float(acc[c][j]);
                }
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty + as);
                if (p.split) {
                    // lo rows publish, hi rows add: xb[doc][row & 63]
                    if (is_lo) {
#pragma unroll
                        for (int j = 0; j < kTsDocs; ++j) xb[j * 64 + (row & 63)] = v[j];
                    }
                    __syncwarp();
                    named_bar_sync(2 + (warp & 1), 64);  // warps (0,2) and (1,3) pair up
                    if (!is_lo) {
#pragma unroll
                        for (int j = 0; j < kTsDocs; ++j) v[j] += xb[j * 64 + row];
                    }
                }
                float mx[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) mx[j] = fmaxf(fmaxf(v[j], v[16 + j]), fmaxf(v[32 + j], v[48 + j]));
#pragma unroll
                for (int st = 8; st >= 1; st >>= 1)
#pragma unroll
                    for (int j = 0; j < st; ++j) mx[j] = fmaxf(mx[j], mx[j + st]);
                if (__ballot_sync(kFullMask, mx[0] >= tau) != 0) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        const int c0 = c * 16;
                        float w[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            w[j] = c == 0 ? v[j] : (c == 1 ? v[16 + j] : (c == 2 ? v[32 + j] : v[48 + j]));
                        float m = w[0];
#pragma unroll
                        for (int j = 1; j < 16; ++j) m = fmaxf(m, w[j]);
                        if (__ballot_sync(kFullMask, m >= tau) != 0) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (w[j] >= tau && c0 + j < ndoc) {
                                    buf_s[cnt * LR + lrow] = w[j];
                                    buf_i[cnt * LR + lrow] = (uint32_t)(doc0 + c0 + j);
                                    ++cnt;
                                }
                            }
                            if (__ballot_sync(kFullMask, cnt > CAP - 16) != 0) {
                                flush();
                                ++n_flush;
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
            for (int c0 = 0; c0 < kTsDocs; c0 += 16) {
                uint32_t acc[16];
                ptx::tmem_ld16(taddr + c0, acc);
                ptx::tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
                if (p.split) {
                    // lo rows publish, hi rows add: xb[doc][row & 63]
                    if (is_lo) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) xb[(c0 + j) * 64 + (row & 63)] = v[j];
                    }
                    __syncwarp();
                    named_bar_sync(2 + (warp & 1), 64);  // warps (0,2) and (1,3) pair up
                    if (!is_lo) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += xb[(c0 + j) * 64 + row];
                    }
                }
                float m = v[0];
#pragma unroll
                for (int j = 1; j < 16; ++j) m = fmaxf(m, v[j]);
                if (__ballot_sync(kFullMask, m >= tau) != 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (v[j] >= tau && c0 + j < ndoc) {
                            buf_s[cnt * LR + lrow] = v[j];
                            buf_i[cnt * LR + lrow] = (uint32_t)(doc0 + c0 + j);
                            ++cnt;
                        }
                    }
                    if (__ballot_sync(kFullMask, cnt > CAP - 16) != 0) flush();
                }
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty + as);
            }  // classic epilogue (QS = false)
        }
        if (tl && lane == 0) {
            unsigned long long *o = p.timeline + (size_t)blockIdx.x * 32;
            o[5] = w_tfull;
            o[6] = ptx::sm_clock() - t_loop;
            o[7] = n_flush;
        }
        flush();
        // 3. publish this row's list
        if (live && !is_lo) {
            float *cs = p.cand_s + (long long)blockIdx.x * p.cand_stride + (long long)(q0 + qrow) * p.k;
            uint32_t *ci = p.cand_i + (long long)blockIdx.x * p.cand_stride + (long long)(q0 + qrow) * p.k;
            if constexpr (KL > 0) {
#pragma unroll
                for (int e = 0; e < KLR; ++e)
                    if (e < p.k) {
                        cs[e] = rs[e];
                        ci[e] = ri[e];
                    }
            } else {
                for (int e = 0; e < p.k; ++e) {
                    cs[e] = lst_s[e * LR + lrow];
                    ci[e] = lst_i[e * LR + lrow];
                }
            }
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (p.timeline != nullptr && tid == 0) p.timeline[(size_t)blockIdx.x * 32 + 15] = ptx::globaltimer_ns();
    if (warp == 4) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, 512);
    }
    if (p.multicast) ptx::cluster_sync_all();
}

}  // namespace vqa
