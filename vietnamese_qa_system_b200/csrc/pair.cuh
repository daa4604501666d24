// pair.cuh -- K2d + K3: the TENSOR-BOUND regime (B >= 256): a CTA PAIR (thread-block cluster of 2 on one TPC)
// drives tcgen05.mma.cta_group::2 -- M = 256 query rows (128 in each CTA's tensor memory), N = 128 documents per
// instruction (64 in each CTA's shared memory), K = 16.
//
// Replaces faiss IndexFlatIP search under txtai ann.search (heavy_ranker.py:98,100) where the pass over the
// documents is a real GEMM: BASELINE configs[1] at B = 1024 and configs[2] at B = 256.
//
// Why: ts.cuh's single-CTA tiles (M = 128 queries x N = 64 documents) are 32-cycle instructions; measured on the B200
// (profiles/r2_ts_b256.txt) the tensor pipe is busy 47 % of the time at B = 256 and the HBM 46 % -- neither bound.
// Per MMA the SM writes the document slice to shared memory once (TMA) and reads it once (B operand): 2 x 64 B/cycle
// at the MMA floor, the whole of the SM's shared-memory bandwidth, and an N = 128 tile does not fit next to the
// query block in tensor memory.  With cta_group::2 each SM stores and reads HALF of every 128-document tile (its
// own 64 rows; the tensor cores of the pair share the halves) and the instruction is 64 cycles long: half the
// shared-memory traffic and half the issue rate per flop, and one HBM pass serves 256 queries with no multicast.
//
// Pair semantics (pinned on the hardware by tools/cta2_probe.cu, profiles/r2_call1.log): A rows 0..127 from CTA 0
// (tensor or shared memory, same offset in both CTAs), 128..255 from CTA 1; B rows 0..N/2-1 from CTA 0's shared
// memory, N/2..N-1 from CTA 1's; D rows 0..127 in CTA 0's tensor memory, 128..255 in CTA 1's, N columns each;
// BOTH CTAs execute tcgen05.alloc.cta_group::2; the leader (rank 0) issues every MMA and signals both CTAs with
// tcgen05.commit.cta_group::2 ... multicast::cluster.
//
// Layout per CTA: tensor memory = [query block: KT x 32 columns][AS accumulator stages x 128 columns];
// shared memory = [KS query blocks x 16 KB (SS-form A)][ring of 8 KB document boxes][barriers][candidate buffers].
// Warps: 0-3 epilogue (thread = query row, register top-k list, as ts.cuh's QS variant), 4 TMA producer (both
// CTAs, each its own 64 documents of every tile), 5 MMA issuer (leader only).
// Screen mode only (storage-precision queries, k + spare <= 32, exact re-scoring in the reduce).
//
// Roofline: tensor pipe (2 * 256 * n_rows * dim flops per launch) above ~230 queries per HBM pass, HBM below.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "mma.cuh"
#include "ptx.cuh"
#include "ts.cuh"

namespace vqa {

constexpr int kPairDocs = 128;   // documents per pair tile (UMMA N); each CTA holds kPairDocs / 2 of them
constexpr int kPairRows = 128;   // query rows per CTA (half of UMMA M = 256)

struct PairParams {
    const float *q;
    long long q_stride;
    int nq;        // queries in this launch (<= 256): CTA rank r of every pair serves queries [128 r, 128 r + 128)
    int k;         // list length inside the scan (k + spare ranks, <= 32)
    long long n_rows;
    int dim;       // multiple of 64, <= 1024
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
    int n_tiles;   // tiles of 128 documents
    int n_stages;
    int kps;
    unsigned long long tma_policy;
    unsigned long long *tau_g;
    uint32_t epoch;
    int ks;        // the LAST ks 64-column blocks of the query block live in shared memory (SS-form MMAs)
    unsigned long long *timeline;  // diagnostic counters as in TsParams, normally nullptr
};

// [align slack][ks query blocks x 16 KB][ring: boxes x 8 KB][barriers 1 KB][candidate buffers 128 rows x 32 x 8 B]
inline size_t pair_smem_bytes_rt(int boxes, int ks) {
    return 1024 + (size_t)ks * kTsQBlockBytes + (size_t)boxes * kTsBoxBytes + 1024 + (size_t)kPairRows * 32 * 8;
}

namespace ptx2 {  // cta_group::2 / cluster-scope forms (not in ptx.cuh: the CPU emulator has no model of CTA pairs)

__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, 256 x N] (+)= A[tmem of both CTAs] * B[smem halves of both CTAs]^T ; leader CTA, ONE thread
__device__ __forceinline__ void umma2_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
        "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive::one on the barrier at this shared-memory offset in every CTA of `mask` once the pair's MMAs retire
__device__ __forceinline__ void umma2_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            ptx::smem_u32(bar)),
        "h"(mask)
        : "memory");
}
// TMA 2-D load into THIS CTA's shared memory whose completion bytes are counted on the barrier at cluster address
// `bar_cluster` -- with .cta_group::2 that barrier may live in the peer CTA (the pair's leader)
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const void *tmap, int32_t c0, int32_t c1,
                                                 uint32_t bar_cluster, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(ptx::smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
// arrive (+ expected transaction bytes) on a barrier given by its cluster address
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// wait with cluster-scope acquire (the arrivals come from the peer CTA); bounded like ptx::mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}\n"
            : "=r"(ok)
            : "r"(ptx::smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (++spins > VQA_SPIN_LIMIT) __trap();
    }
}

}  // namespace ptx2

// PAIR = true: launched as clusters of 2 (cta_group::2, 256 queries per pair).  PAIR = false: the same 128-document
// tiles on ONE CTA (cta_group::1, M = 128 queries, N = 128 documents: a 64-cycle instruction that reads its 4 KB of A
// from tensor memory at the 64 B/cycle the TMEM read port gives -- ts.cuh's N = 64 tiles need that bandwidth twice
// over, which is what holds B = 64..128 at ~58 cycles per 32-cycle MMA, profiles/r2_ts_waits_10m.json); the CTA
// fetches both 64-row halves of a tile itself.
template <bool DOC_BF16, int KL, bool PAIR>
__global__ void __launch_bounds__(kMmaThreads, 1)
ts_pair_topk_kernel(const __grid_constant__ CUtensorMap tmap_docs, const PairParams p) {
    static_assert(KL == 16 || KL == 32, "register lists only");
    constexpr int HB = PAIR ? 1 : 2;   // 64-row document boxes this CTA fetches per 64-column block of a tile
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int KB = p.dim / kBlockK;
    const int S = p.n_stages;
    const int KPS = p.kps;
    const int KG = KB / KPS;
    const uint32_t stage_bytes = (uint32_t)KPS * HB * kTsBoxBytes;  // per CTA: KPS x HB boxes of 64 documents x 64 columns
    const int KSB = p.ks;
    const int KT = KB - KSB;
    const int ACOLS = KT * (kBlockK / 2);
    const int AS = (512 - ACOLS) / kPairDocs;                   // accumulator stages of 128 columns
    unsigned char *q_smem = smem;
    unsigned char *a_smem = smem + (size_t)KSB * kTsQBlockBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(a_smem + (size_t)S * stage_bytes);
    uint64_t *full = bars;                          // leader: 2 producer arrivals + both CTAs' bytes
    uint64_t *empty = bars + kMaxStages;            // each CTA: 1 arrival (leader's multicast commit)
    uint64_t *tfull = bars + 2 * kMaxStages;        // each CTA: 1 arrival (leader's multicast commit)
    uint64_t *tempty = tfull + kMaxAccStages;       // leader: 8 arrivals (4 epilogue warps of each CTA)
    uint64_t *qready = tempty + kMaxAccStages;      // leader: 8 arrivals (query block staged in both CTAs)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(qready + 1);
    unsigned char *lists = reinterpret_cast<unsigned char *>(bars) + 1024;
    constexpr int CAP = 32;
    float *buf_s = reinterpret_cast<float *>(lists);            // [CAP][128] candidate buffers, entry-major
    uint32_t *buf_i = reinterpret_cast<uint32_t *>(buf_s + CAP * kPairRows);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);
    const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;   // 0 = leader
    const int pair0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_pairs = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int q0 = (int)rank * kPairRows;
    const int nq = p.nq - q0 < kPairRows ? (p.nq - q0 > 0 ? p.nq - q0 : 0) : kPairRows;
    // M = 256 (pair) / 128, N = 128, K-major A and B, fp32 accumulate
    const uint32_t idesc = (1u << 4) | ((DOC_BF16 ? 1u : 0u) << 7) | ((DOC_BF16 ? 1u : 0u) << 10) |
                           ((uint32_t)(kPairDocs >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    constexpr uint32_t NARR = PAIR ? 8u : 4u;   // epilogue warps that report to the MMA warp

    if (p.timeline != nullptr && tid == 0) p.timeline[(size_t)blockIdx.x * 32] = ptx::globaltimer_ns();
    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&tmap_docs);
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(full + s, PAIR ? 2 : 1);
            ptx::mbar_init(empty + s, 1);
        }
        for (int a = 0; a < AS; ++a) {
            ptx::mbar_init(tfull + a, 1);
            ptx::mbar_init(tempty + a, NARR);
        }
        ptx::mbar_init(qready, NARR);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    if constexpr (PAIR) ptx::cluster_sync_all();   // the peer signals these barriers: they must exist pair-wide first
    if (warp == 4) {
        if constexpr (PAIR) {
            ptx2::tmem_alloc2(tmem_slot, 512);
            ptx2::tmem_relinquish2();
        } else {
            ptx::tmem_alloc(tmem_slot, 512);
            ptx::tmem_relinquish();
        }
        ptx::tc_fence_before_sync();
    }
    grid_launch_dependents();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(kFullMask, *reinterpret_cast<volatile uint32_t *>(tmem_slot), 0);

    if (warp == 4) {
        // ===== TMA producer (both CTAs): this CTA's 64 documents of every pair tile; bytes counted on the LEADER's
        // full barrier, the stage is free again when the leader's commit arrives on OUR empty barrier =====
        uint32_t it = 0;
        const bool tl = p.timeline != nullptr;
        unsigned long long w_empty = 0, t_loop = tl ? ptx::sm_clock() : 0;
        for (int tile = pair0; tile < p.n_tiles; tile += n_pairs) {
            for (int kg = 0; kg < KG; ++kg, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                const unsigned long long t_w = tl ? ptx::sm_clock() : 0;
                ptx::mbar_wait(empty + s, ph ^ 1);
                if (tl) w_empty += ptx::sm_clock() - t_w;
                if (ptx::elect_one()) {
                    if constexpr (PAIR) {
                        const uint32_t lead_full = ptx2::mapa_rank(ptx::smem_u32(full + s), 0);
                        ptx2::mbar_arrive_expect_tx_cluster(lead_full, stage_bytes);
                        for (int j = 0; j < KPS; ++j)
                            ptx2::tma_load_2d_pair(a_smem + (size_t)s * stage_bytes + (size_t)j * kTsBoxBytes, &tmap_docs,
                                                   (kg * KPS + j) * kBlockK, tile * kPairDocs + (int)rank * (kPairDocs / 2),
                                                   lead_full, p.tma_policy);
                    } else {
                        ptx::mbar_arrive_expect_tx(full + s, stage_bytes);
                        for (int j = 0; j < KPS; ++j)
#pragma unroll
                            for (int hb = 0; hb < 2; ++hb)   // rows 0..63 and 64..127 of the tile, back to back: one 128-row B operand
                                ptx::tma_load_2d(a_smem + (size_t)s * stage_bytes + (size_t)(j * 2 + hb) * kTsBoxBytes, &tmap_docs,
                                                 (kg * KPS + j) * kBlockK, tile * kPairDocs + hb * (kPairDocs / 2), full + s,
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
        if (rank == 0) {
            // ===== MMA issuer (leader only): waits for both CTAs' query blocks, then D = Q * docs^T for the pair =====
            if constexpr (PAIR) ptx2::mbar_wait_cluster(qready, 0);
            else ptx::mbar_wait(qready, 0);
            ptx::tc_fence_after_sync();
            uint32_t it = 0, lt = 0;
            const uint32_t ring = ptx::smem_u32(a_smem);
            const uint32_t qs_base = ptx::smem_u32(q_smem);
            const bool tl = p.timeline != nullptr;
            unsigned long long w_full = 0, w_tempty = 0, t_loop = tl ? ptx::sm_clock() : 0;
            for (int tile = pair0; tile < p.n_tiles; tile += n_pairs, ++lt) {
                const int as = lt % AS;
                const uint32_t aph = (lt / AS) & 1;
                unsigned long long t_w = tl ? ptx::sm_clock() : 0;
                if constexpr (PAIR) ptx2::mbar_wait_cluster(tempty + as, aph ^ 1);
                else ptx::mbar_wait(tempty + as, aph ^ 1);
                if (tl) w_tempty += ptx::sm_clock() - t_w;
                ptx::tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(ACOLS + as * kPairDocs);
                for (int kg = 0; kg < KG; ++kg, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    t_w = tl ? ptx::sm_clock() : 0;
                    if constexpr (PAIR) ptx2::mbar_wait_cluster(full + s, ph);
                    else ptx::mbar_wait(full + s, ph);
                    if (tl) w_full += ptx::sm_clock() - t_w;
                    ptx::tc_fence_after_sync();
                    if (ptx::elect_one()) {
                        for (int j = 0; j < KPS; ++j) {
                            const int kb = kg * KPS + j;
                            const uint64_t db0 = ptx::umma_desc_k_sw128(ring + (uint32_t)s * stage_bytes + (uint32_t)(j * HB) * kTsBoxBytes);
                            if (kb >= KT) {  // this block of the queries is in shared memory (both CTAs, same offset)
                                const uint64_t da0 = ptx::umma_desc_k_sw128(qs_base + (uint32_t)(kb - KT) * kTsQBlockBytes);
#pragma unroll
                                for (int k4 = 0; k4 < kBlockK / 16; ++k4) {
                                    if constexpr (PAIR) ptx2::umma2_ss(d_tmem, da0 + (uint64_t)(k4 * 2), db0 + (uint64_t)(k4 * 2), idesc, (kb | k4) != 0 ? 1u : 0u);
                                    else ptx::umma_f16(d_tmem, da0 + (uint64_t)(k4 * 2), db0 + (uint64_t)(k4 * 2), idesc, (kb | k4) != 0 ? 1u : 0u);
                                }
                                continue;
                            }
#pragma unroll
                            for (int k4 = 0; k4 < kBlockK / 16; ++k4) {
                                if constexpr (PAIR) ptx2::umma2_ts(d_tmem, tmem_base + (uint32_t)(kb * 32 + k4 * 8), db0 + (uint64_t)(k4 * 2), idesc, (kb | k4) != 0 ? 1u : 0u);
                                else umma_f16_ts(d_tmem, tmem_base + (uint32_t)(kb * 32 + k4 * 8), db0 + (uint64_t)(k4 * 2), idesc, (kb | k4) != 0 ? 1u : 0u);
                            }
                        }
                        if constexpr (PAIR) {
                            ptx2::umma2_commit_mc(empty + s, (uint16_t)0x3);                    // stage free in both CTAs
                            if (kg == KG - 1) ptx2::umma2_commit_mc(tfull + as, (uint16_t)0x3);  // accumulator ready in both
                        } else {
                            ptx::umma_commit(empty + s);
                            if (kg == KG - 1) ptx::umma_commit(tfull + as);
                        }
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
        }
    } else {
        // ===== warps 0-3 (both CTAs): thread = query row =====
        const int row = warp * 32 + lane;            // TMEM lane = query row of this CTA
        const bool live = row < nq;
        const uint32_t lead_qready = PAIR ? ptx2::mapa_rank(ptx::smem_u32(qready), 0) : 0u;
        // 1. this row of the query block -> tensor memory (first KT blocks) and shared memory (last KSB blocks)
        {
            const float *src = p.q + (long long)(q0 + (live ? row : 0)) * p.q_stride;
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
            constexpr int QG = 4;
            for (int c0 = 0; c0 < ACOLS; c0 += 16 * QG) {
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
                        if (DOC_BF16) {
                            w[2 * jj] = pack_bf16x2(v[g][jj].x, v[g][jj].y);
                            w[2 * jj + 1] = pack_bf16x2(v[g][jj].z, v[g][jj].w);
                        } else {
                            w[2 * jj] = pack_f16x2(v[g][jj].x, v[g][jj].y);
                            w[2 * jj + 1] = pack_f16x2(v[g][jj].z, v[g][jj].w);
                        }
                    }
                    tmem_st16(trow + (uint32_t)(c0 + g * 16), w);
                }
            }
            tmem_st_wait();
            for (int kb = KT; kb < KB; ++kb) {
                unsigned char *tile = q_smem + (size_t)(kb - KT) * kTsQBlockBytes;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (live) {
                        v0 = *reinterpret_cast<const float4 *>(src + kb * kBlockK + c * 8);
                        v1 = *reinterpret_cast<const float4 *>(src + kb * kBlockK + c * 8 + 4);
                    }
                    uint4 w;
                    if (DOC_BF16) w = make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
                    else w = make_uint4(pack_f16x2(v0.x, v0.y), pack_f16x2(v0.z, v0.w), pack_f16x2(v1.x, v1.y), pack_f16x2(v1.z, v1.w));
                    *reinterpret_cast<uint4 *>(tile + row * 128 + ((c ^ (row & 7)) << 4)) = w;
                }
            }
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {   // 4 warps (x 2 CTAs) -> the (leader's) MMA warp may start
                if constexpr (PAIR) ptx2::mbar_arrive_cluster(lead_qready);
                else ptx::mbar_arrive(qready);
            }
        }
        // 2. private register list + threshold (ts.cuh, QS epilogue)
        float tau = live ? neg_inf() : __int_as_float(0x7f800000);
        unsigned long long *tg = (p.tau_g != nullptr && live) ? p.tau_g + q0 + row : nullptr;
        float rs[KL];
        uint32_t ri[KL];
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < KL; ++e) {
            rs[e] = neg_inf();
            ri[e] = invalid_id<uint32_t>();
        }
        auto flush = [&]() {
            const int wmax = __reduce_max_sync(kFullMask, cnt);
            for (int b = 0; b < wmax; ++b) {
                float cs = neg_inf();
                uint32_t ci = invalid_id<uint32_t>();
                if (b < cnt) {
                    cs = buf_s[b * kPairRows + row];
                    ci = buf_i[b * kPairRows + row];
                }
                bool bef[KL];
#pragma unroll
                for (int e = 0; e < KL; ++e) bef[e] = ranks_before<uint32_t>(cs, ci, rs[e], ri[e]);
#pragma unroll
                for (int e = KL - 1; e >= 0; --e) {
                    const bool up = e > 0 ? bef[e - 1] : false;
                    rs[e] = up ? rs[e > 0 ? e - 1 : 0] : (bef[e] ? cs : rs[e]);
                    ri[e] = up ? ri[e > 0 ? e - 1 : 0] : (bef[e] ? ci : ri[e]);
                }
            }
            cnt = 0;
            float ts = __int_as_float(0x7f800000);
            int filled = 0;
#pragma unroll
            for (int e = 0; e < KL; ++e) {
                ts = fminf(ts, e < p.k ? rs[e] : __int_as_float(0x7f800000));
                filled += (e < p.k && ri[e] != invalid_id<uint32_t>()) ? 1 : 0;
            }
            if (filled == p.k && ts > tau) {
                tau = ts;
                if (tg != nullptr) atomicMax(tg, tau_encode(ts, p.epoch));
            }
        };

        const uint32_t lead_tempty0 = PAIR ? ptx2::mapa_rank(ptx::smem_u32(tempty), 0) : 0u;
        uint32_t lt = 0;
        const bool tl = p.timeline != nullptr && warp == 0;
        unsigned long long w_tfull = 0, n_flush = 0, t_loop = tl ? ptx::sm_clock() : 0;
        for (int tile = pair0; tile < p.n_tiles; tile += n_pairs, ++lt) {
            const int as = lt % AS;
            const uint32_t aph = (lt / AS) & 1;
            unsigned long long graw = 0;
            if (tg != nullptr) graw = ld_volatile_u64(tg);
            const unsigned long long t_w = tl ? ptx::sm_clock() : 0;
            ptx::mbar_wait(tfull + as, aph);
            if (tl) w_tfull += ptx::sm_clock() - t_w;
            ptx::tc_fence_after_sync();
            if (tg != nullptr) tau = fmaxf(tau, tau_decode(graw, p.epoch));
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ACOLS + as * kPairDocs);
            const long long doc0 = (long long)tile * kPairDocs;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float v[64];
                {
                    uint32_t acc[4][16];
#pragma unroll
                    for (int c = 0; c < 4; ++c) ptx::tmem_ld16(taddr + half * 64 + c * 16, acc[c]);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[c * 16 + j] = __uint_as_float(acc[c][j]);
                }
                if (half == 1) {  // all 128 scores of the row are in registers: the accumulator goes back to the MMA warp
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (PAIR) ptx2::mbar_arrive_cluster(lead_tempty0 + (uint32_t)as * 8u);
                        else ptx::mbar_arrive(tempty + as);
                    }
                }
                const long long hdoc0 = doc0 + half * 64;
                const int ndoc = p.n_rows - hdoc0 < 64 ? (p.n_rows - hdoc0 > 0 ? (int)(p.n_rows - hdoc0) : 0) : 64;
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
                                    buf_s[cnt * kPairRows + row] = w[j];
                                    buf_i[cnt * kPairRows + row] = (uint32_t)(hdoc0 + c0 + j);
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
            }
        }
        if (tl && lane == 0) {
            unsigned long long *o = p.timeline + (size_t)blockIdx.x * 32;
            o[5] = w_tfull;
            o[6] = ptx::sm_clock() - t_loop;
            o[7] = n_flush;
            if (rank != 0) o[8] = lt;
        }
        flush();
        // 3. publish this row's list: list index = blockIdx.x, so query q appears in the lists l with l % 2 == q / 128
        if (live) {
            float *cs = p.cand_s + (long long)blockIdx.x * p.cand_stride + (long long)(q0 + row) * p.k;
            uint32_t *ci = p.cand_i + (long long)blockIdx.x * p.cand_stride + (long long)(q0 + row) * p.k;
#pragma unroll
            for (int e = 0; e < KL; ++e)
                if (e < p.k) {
                    cs[e] = rs[e];
                    ci[e] = ri[e];
                }
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if constexpr (PAIR) ptx::cluster_sync_all();   // the leader's MMAs read the peer's shared / tensor memory: nobody leaves early
    if (p.timeline != nullptr && tid == 0) p.timeline[(size_t)blockIdx.x * 32 + 15] = ptx::globaltimer_ns();
    if (warp == 4) {
        ptx::tc_fence_after_sync();
        if constexpr (PAIR) ptx2::tmem_dealloc2(tmem_base, 512);
        else ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace vqa
