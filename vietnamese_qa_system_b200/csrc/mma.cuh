// mma.cuh -- K2b + K3: tcgen05/TMEM tensor-core similarity tiles fed by TMA, with
// the top-k selection fused into the TMEM epilogue.
//
// Replaces faiss IndexFlatIP search under txtai ann.search (heavy_ranker.py:98,100)
// for LARGER query batches, where the pass over the documents is a real
// [n_rows x dim] x [dim x B] contraction.
//
// Persistent, warp-specialised CTA (one per SM, 192 threads):
//   warps 0-3  epilogue: tcgen05.ld the 128 x N fp32 score tile out of TMEM
//              (thread = one document row), compare against the per-query
//              thresholds, rare insert into the CTA-shared top-k lists;
//   warp 4     TMA producer: streams 128-row x 64-col (16 KB, 128B-swizzled)
//              document blocks through an S-stage mbarrier ring; owns TMEM alloc;
//   warp 5     MMA issuer: one elected thread issues tcgen05.mma (M=128 docs,
//              N=query columns, K=16) from shared-memory descriptors into one of
//              AS TMEM accumulator stages, tcgen05.commit frees smem / hands the
//              accumulator to the epilogue.
// The queries stay resident in shared memory for the whole kernel, converted
// on the fly from fp32 to the storage type as a hi + lo pair (two MMA columns
// per query, added in the epilogue) so that the only rounding left is the
// documents' own storage rounding -- this is what keeps recall@k >= 0.999
// against the fp32 verify mode.  (Column pairs rather than a second MMA into the
// same columns: the issue rate of small tcgen05.mma instructions, not the tensor
// pipe, is what limits this kernel, so fewer, wider MMAs win.)
// The [B, n_rows] score matrix never leaves the SM.
//
// Roofline: HBM up to B ~ 200 (algorithmic bytes = n_rows*dim*2 per launch),
// tensor pipe beyond.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace vqa {

struct MmaParams {
    const float *q;
    long long q_stride;
    int nq;  // queries in this LAUNCH: n_groups chunks of NCOL / 2 (hi and lo column per query); the last may be short
    int n_groups;  // query chunks processed side by side: CTA c works on chunk c % n_groups, tile stream
                   // c / n_groups (one HBM pass serves n_groups * NCOL/2 queries)
    int multicast; // 1: launched as clusters of n_groups CTAs; every document box is fetched ONCE per
                   // cluster -- each CTA issues a 128/n_groups-row slice of it as a TMA multicast into all
                   // the cluster's shared memories -- so L2->SM traffic does not grow with n_groups.
                   // 0: plain CTAs, the chunks' CTAs re-read the tiles through L2.
    int k;
    long long n_rows;
    int dim;  // multiple of 64
    float lo_scale;      // q_lo is stored as (q - q_hi) * lo_scale ...
    float lo_inv_scale;  // ... and score = acc_hi + acc_lo * lo_inv_scale
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
    int n_tiles;
    int n_stages;  // smem ring depth
    int kps;       // k-blocks (16 KB TMA boxes) per ring stage
    unsigned long long tma_policy;  // L2 cache hint for the document stream
    // GPU-wide per-query thresholds shared by all CTAs: (epoch << 32) | ordered(score of some list's
    // k-th best).  A document scoring below any list's k-th best cannot be in the global top-k.
    unsigned long long *tau_g;  // [nq] for this pass, or nullptr
    uint32_t epoch;
    // Warm-up seed (register-list path): [nq][kSlotStride] slots, same encoding as tau_g.  While a list is empty
    // nothing can be rejected, so tile 0 used to cost one 32-wide bitonic merge per query and warp -- a ~2000-cycle
    // dependent chain each, 40 us at B = 32, 75 us over the first eight tiles (profiles/r2_timeline_shard.json).
    // Instead, for its FIRST tile a warp only publishes, per query, the best score among its 32 rows into slot
    // (global warp % k) -- five shuffle steps -- and holds the accumulator stage; before its second tile it reads the
    // query's k slots back.  They hold scores of k DISTINCT documents (lists are disjoint), so their minimum is a
    // lower bound of the global k-th best, and it is already tight: every slot is the best of ~1900 rows.  Then
    // tile 0 is filtered against that bound like any later tile (a handful of single inserts).  nullptr: off.
    unsigned long long *slot_g;
    // Dynamic tile schedule (launches without clusters): one 64-bit counter in the workspace, (epoch << 32) | next
    // tile.  The producer warp of every CTA takes tiles from it instead of the static round-robin share, so an SM
    // that streams slower than its neighbours (the per-CTA spread was 195..270 us of a 285 us launch at B = 1,
    // profiles/r2_timeline_shard.json) simply takes fewer tiles and all CTAs finish together.  nullptr: static.
    unsigned long long *tile_ctr;
    // Diagnostic (vqa_debug_timeline): [gridDim.x][kTimelineSlots] per-CTA stamps -- %globaltimer in slots 0..15,
    // clock64 in 16..31 -- of entry, first/last TMA issue, first MMA / last commit, queries staged, tiles
    // 0, 1, 3, 7, 15, 31, 63 and the last one leaving the epilogue, and exit.  nullptr (always, in normal use): off.
    unsigned long long *timeline;
};

constexpr int kSlotStride = 32;  // seed slots reserved per query (k <= 32 on the register-list path)
constexpr int kTimelineSlots = 32;
__device__ __forceinline__ void timeline_stamp(unsigned long long *tl, int slot) {
    if (tl != nullptr && (threadIdx.x & 31) == 0) {
        tl[(size_t)blockIdx.x * kTimelineSlots + slot] = ptx::globaltimer_ns();
        tl[(size_t)blockIdx.x * kTimelineSlots + 16 + slot] = ptx::sm_clock();
    }
}

// order-preserving float -> uint32 (larger float <=> larger uint), tagged with the search epoch
__device__ __forceinline__ unsigned long long tau_encode(float f, uint32_t epoch) {
    const uint32_t u = __float_as_uint(f);
    const uint32_t ord = (u >> 31) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)epoch << 32) | ord;
}
__device__ __forceinline__ float tau_decode(unsigned long long v, uint32_t epoch) {
    if ((uint32_t)(v >> 32) != epoch) return neg_inf();  // stale / uninitialised slot
    const uint32_t ord = (uint32_t)v;
    return __uint_as_float((ord >> 31) ? (ord & 0x7fffffffu) : ~ord);
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- tile schedule -------------------------------------------------------------------------------
constexpr int kTileRing = 32;  // published tile ids in flight per CTA (producer runs < 16 tiles ahead of the epilogue)

// Every CTA's first tile is its static one (tile = CTA index, no atomic in front of the first TMA); later tiles are
// tickets of one counter, (epoch << 32) | tickets taken, tile = gridDim.x + ticket.  Tickets are plain atomicAdds issued
// one tile ahead, so their ~1 us round trip hides behind the TMA issue loop.  A counter left at zero by the previous
// launch (or holding an older epoch) is brought to (epoch, 0) by an atomicMax every CTA issues first; a word that is
// neither (a workspace that was never zeroed) is repaired through a CAS -- correct, but a one-off ~0.3 ms storm, which
// is why vqa.h asks for a zeroed workspace.  Exactly gridDim.x tickets are >= the number of dynamic tiles (every CTA
// stops at its first); whoever draws the LAST of them zeroes the counter for the next launch / graph replay.
__device__ __forceinline__ int tile_from_ticket(unsigned long long *ctr, unsigned long long old, uint32_t epoch, int n_tiles) {
    uint32_t t = (uint32_t)old;
    if ((uint32_t)(old >> 32) != epoch) {  // garbage word: take a ticket through a CAS that also installs the epoch
        unsigned long long cur = ld_volatile_u64(ctr);
        while (true) {
            const bool fresh = (uint32_t)(cur >> 32) != epoch;
            t = fresh ? 0u : (uint32_t)cur;
            const unsigned long long nv = ((unsigned long long)epoch << 32) | (unsigned long long)(t + 1);
            const unsigned long long seen = atomicCAS(ctr, cur, nv);
            if (seen == cur) break;
            cur = seen;
        }
    }
    if (t == (uint32_t)n_tiles - 1u) *reinterpret_cast<volatile unsigned long long *>(ctr) = 0ull;
    const long long tile = (long long)gridDim.x + t;
    return tile < n_tiles ? (int)tile : -1;
}

// tile `lt` of this CTA as seen by a consumer role (MMA issuer, epilogue): the static share, or what the producer
// warp published (spin on shared memory; the producer is always ahead except at the very start)
__device__ __forceinline__ int consumer_tile(const MmaParams &p, const volatile uint32_t *pub, const volatile int *tile_q,
                                             uint32_t lt, int stream0, int n_streams) {
    if (p.tile_ctr == nullptr) {
        const long long t = stream0 + (long long)lt * n_streams;
        return t < p.n_tiles ? (int)t : -1;
    }
    while (*pub <= lt) __nanosleep(20);
    __threadfence_block();
    return tile_q[lt % kTileRing];
}

template <int NCOL>
__host__ __device__ constexpr int mma_acc_stages() {
    return (512 / NCOL) < kMaxAccStages ? (512 / NCOL) : kMaxAccStages;
}

// shared memory: [Q: dim/64 tiles of NCOL x 128 B][A ring: stages x 16 KB][barriers][lists]
// (size: mma_smem_bytes_rt in consts.h)

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

// hi/lo split of 8 consecutive fp32 values into two 16-byte chunks
template <bool BF16>
__device__ __forceinline__ void split8(const float (&x)[8], float lo_scale, uint4 &hi, uint4 &lo) {
    float h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if constexpr (BF16) {
            h[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
        } else {
            h[j] = __half2float(__float2half_rn(x[j]));
        }
        l[j] = (x[j] - h[j]) * lo_scale;  // the subtraction is exact in fp32
    }
    if constexpr (BF16) {
        hi = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]),
                        pack_bf16x2(h[6], h[7]));
        lo = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]),
                        pack_bf16x2(l[6], l[7]));
    } else {
        hi = make_uint4(pack_f16x2(h[0], h[1]), pack_f16x2(h[2], h[3]), pack_f16x2(h[4], h[5]),
                        pack_f16x2(h[6], h[7]));
        lo = make_uint4(pack_f16x2(l[0], l[1]), pack_f16x2(l[2], l[3]), pack_f16x2(l[4], l[5]),
                        pack_f16x2(l[6], l[7]));
    }
}

// Warp-wide bulk insert of up to 32 candidates (one per lane; lanes without a
// candidate pass (-inf, invalid)) into list q.  Only for lists with k <= 32.
// Bitonic sort of the candidates, bitonic merge with the (sorted) list.
__device__ __forceinline__ void bitonic_step(float &s, uint32_t &i, int stride, bool keep_before) {
    const float ps = __shfl_xor_sync(kFullMask, s, stride);
    const uint32_t pi = __shfl_xor_sync(kFullMask, i, stride);
    const bool mine_before = ranks_before<uint32_t>(s, i, ps, pi);
    if (mine_before != keep_before) {
        s = ps;
        i = pi;
    }
}

static __device__ __noinline__ void list_insert_bulk32(ListView<uint32_t> L, int q, float cs, uint32_t ci) {
    const int lane = threadIdx.x & 31;
    // sort candidates descending across lanes
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const bool desc = (lane & size) == 0 || size == 32;
            const bool lower = (lane & stride) == 0;
            bitonic_step(cs, ci, stride, lower == desc);
        }
    }
    volatile float *ls = L.s + (size_t)q * L.kcap;
    volatile uint32_t *li = L.i + (size_t)q * L.kcap;
    if (lane == 0) {
        while (atomicCAS(L.lock + q, 0, 1) != 0) __nanosleep(32);
    }
    __syncwarp();
    __threadfence_block();
    // reversed list against sorted candidates -> bitonic sequence holding the top 32 of the union
    float es = ls[31 - lane];
    uint32_t ei = li[31 - lane];
    if (ranks_before<uint32_t>(es, ei, cs, ci)) {
        cs = es;
        ci = ei;
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) bitonic_step(cs, ci, stride, (lane & stride) == 0);
    if (lane < L.k) {
        ls[lane] = cs;
        li[lane] = ci;
        if (lane == L.k - 1) *(volatile float *)(L.tau + q) = cs;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(L.lock + q, 0);
}

// ---- register-resident per-warp lists (k <= 32, <= 32 queries per pass) ---------------------
// Lane e of the warp holds rank e of a sorted (descending) 32-entry list: no shared memory, no
// locks, no fences on the (rare) insert path.
struct Entry {
    float s;
    uint32_t i;
};

static __device__ __noinline__ Entry reglist_insert_one(float ls, uint32_t li, float cs, uint32_t ci) {
    const int lane = threadIdx.x & 31;
    const int pos = __popc(__ballot_sync(kFullMask, ranks_before<uint32_t>(ls, li, cs, ci)));
    const float ups = __shfl_up_sync(kFullMask, ls, 1);
    const uint32_t upi = __shfl_up_sync(kFullMask, li, 1);
    Entry e;
    e.s = lane < pos ? ls : (lane == pos ? cs : ups);
    e.i = lane < pos ? li : (lane == pos ? ci : upi);
    return e;
}

// merge up to 32 unsorted candidates (one per lane, (-inf, invalid) where none) into the list
static __device__ __noinline__ Entry reglist_merge32(float ls, uint32_t li, float cs, uint32_t ci, bool cand_sorted) {
    const int lane = threadIdx.x & 31;
    if (!cand_sorted) {
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const bool desc = (lane & size) == 0 || size == 32;
                bitonic_step(cs, ci, stride, ((lane & stride) == 0) == desc);
            }
        }
    }
    const float rs = __shfl_sync(kFullMask, ls, 31 - lane);
    const uint32_t ri = __shfl_sync(kFullMask, li, 31 - lane);
    if (ranks_before<uint32_t>(rs, ri, cs, ci)) {
        cs = rs;
        ci = ri;
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) bitonic_step(cs, ci, stride, (lane & stride) == 0);
    Entry e;
    e.s = cs;
    e.i = ci;
    return e;
}

// NB independent merges of 32 candidates (one per lane) into NB register lists, interleaved step by step.  One merge is
// a chain of ~20 dependent shuffle+compare steps (~2000 cycles with a single warp per scheduler); while a kernel
// warms up EVERY query of a tile needs one (nothing beats an empty list), and run one after the other they cost the
// headline kernel ~40 us for its first tile and ~75 us over its first eight (profiles/r2_timeline_shard.json).
// Interleaving NB = 8 chains keeps the shuffle pipe busy instead: same instructions, an eighth of the latency.
// Lists ls/li are the epilogue's per-query arrays, `base` the first query of the batch (a constant after unrolling).
template <int NB, int RQ>
__device__ __forceinline__ void reglist_merge32_batch(float (&ls)[RQ], uint32_t (&li)[RQ], int base, float (&cs)[NB],
                                                      uint32_t (&ci)[NB], bool cand_sorted) {
    const int lane = threadIdx.x & 31;
    if (!cand_sorted) {
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const bool keep = ((lane & stride) == 0) == ((lane & size) == 0 || size == 32);
#pragma unroll
                for (int b = 0; b < NB; ++b) bitonic_step(cs[b], ci[b], stride, keep);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const float rs = __shfl_sync(kFullMask, ls[base + b], 31 - lane);
        const uint32_t ri = __shfl_sync(kFullMask, li[base + b], 31 - lane);
        if (ranks_before<uint32_t>(rs, ri, cs[b], ci[b])) {
            cs[b] = rs;
            ci[b] = ri;
        }
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
#pragma unroll
        for (int b = 0; b < NB; ++b) bitonic_step(cs[b], ci[b], stride, (lane & stride) == 0);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ls[base + b] = cs[b];
        li[base + b] = ci[b];
    }
}

// one 16-column group of this thread's document row.  SPLIT: score = acc_hi + acc_lo * lo_inv_scale
// (column j = q_hi[j], column NQ + j = q_lo[j]); otherwise one column per query.
template <int NCOL, bool SPLIT>
__device__ __forceinline__ void load_scores16(uint32_t taddr, int c0, float lo_inv_scale, float (&v)[16]) {
    if constexpr (!SPLIT) {
        uint32_t acc[16];
        ptx::tmem_ld16(taddr + c0, acc);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
    } else {
        constexpr int NQ = NCOL / 2;
        if constexpr (NQ >= 16) {
            uint32_t hi[16], lo[16];
            ptx::tmem_ld16(taddr + c0, hi);
            ptx::tmem_ld16(taddr + NQ + c0, lo);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                v[j] = __fmaf_rn(__uint_as_float(lo[j]), lo_inv_scale, __uint_as_float(hi[j]));
        } else {  // NCOL == 16: one load brings hi (cols 0..7) and lo (cols 8..15)
            uint32_t hl[16];
            ptx::tmem_ld16(taddr, hl);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = __fmaf_rn(__uint_as_float(hl[8 + j]), lo_inv_scale, __uint_as_float(hl[j]));
                v[8 + j] = neg_inf();
            }
        }
    }
}

// SPLIT: hi + lo column per query (NCOL / 2 queries per CTA, scores good to fp32 rounding).
// !SPLIT: one storage-precision column per query (NCOL queries per CTA): a SCREEN whose k + spare best
// candidates are re-scored exactly by the reduce kernel (screen-then-rescore, k + spare <= 32).
template <bool BF16, int NCOL, bool SPLIT>
__global__ void __launch_bounds__(kMmaThreads, 1)
mma_topk_kernel(const __grid_constant__ CUtensorMap tmap_docs, const MmaParams p) {
    constexpr int NQ = SPLIT ? NCOL / 2 : NCOL;  // queries per CTA
    constexpr int AS = mma_acc_stages<NCOL>();
    constexpr uint32_t TMEM_COLS = AS * NCOL;    // power of two, 128..512
    static_assert(NCOL % 16 == 0 && NCOL >= 16 && NCOL <= 256, "MMA N");
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32 && TMEM_COLS <= 512, "TMEM cols");
    constexpr uint32_t IDESC = ptx::umma_idesc_f16(kTileRows, NCOL, BF16);

    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int KB = p.dim / kBlockK;
    const int S = p.n_stages;
    const int KPS = p.kps;                 // KB % KPS == 0
    const int KG = KB / KPS;               // ring stages consumed per tile
    const uint32_t stage_bytes = (uint32_t)KPS * kStageBytes;
    unsigned char *q_smem = smem;                                   // KB tiles of NCOL*128 B
    unsigned char *a_smem = q_smem + (size_t)KB * NCOL * 128;       // S stages of KPS x 16 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(a_smem + (size_t)S * stage_bytes);
    uint64_t *full = bars;                       // [kMaxStages]
    uint64_t *empty = bars + kMaxStages;         // [kMaxStages]
    uint64_t *tfull = bars + 2 * kMaxStages;     // [kMaxAccStages]
    uint64_t *tempty = tfull + kMaxAccStages;    // [kMaxAccStages]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + kMaxAccStages);
    volatile int *tile_q = reinterpret_cast<volatile int *>(reinterpret_cast<unsigned char *>(bars) + 512);  // [kTileRing]
    volatile uint32_t *tile_pub = reinterpret_cast<volatile uint32_t *>(reinterpret_cast<unsigned char *>(bars) + 512 +
                                                                         kTileRing * 4);
    ListView<uint32_t> L = list_carve<uint32_t>(reinterpret_cast<unsigned char *>(bars) + 1024, NQ, p.k);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);  // warp-uniform by construction
    const int grp = blockIdx.x % p.n_groups;               // query chunk of this CTA
    const int stream0 = blockIdx.x / p.n_groups;           // first document tile of this CTA
    const int n_streams = gridDim.x / p.n_groups;
    const int q0 = grp * NQ;                               // first query of the chunk
    const int nq = p.nq - q0 < NQ ? (p.nq - q0 > 0 ? p.nq - q0 : 0) : NQ;  // queries in this CTA's chunk
    const float *qsrc = p.q + (long long)q0 * p.q_stride;
    unsigned long long *tau_g = p.tau_g != nullptr ? p.tau_g + q0 : nullptr;

    // ---- one-time setup ------------------------------------------------------
    // Warp 4 (TMA producer) sets up the barriers and TMEM, then starts streaming documents at once;
    // the other five warps stage the queries meanwhile and meet it at named barrier 1.
    uint32_t tmem_base = 0;
    const uint16_t cta_mask = (uint16_t)((1u << p.n_groups) - 1u);
    const int slice_rows = kTileRows / (p.multicast ? p.n_groups : 1);  // rows of each box this CTA fetches
    if (warp == 0) timeline_stamp(p.timeline, 0);
    if (warp == 4 && lane == 0) {
        ptx::prefetch_tmap(&tmap_docs);
        *tile_pub = 0;
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(full + s, 1);
            ptx::mbar_init(empty + s, p.multicast ? p.n_groups : 1);  // every consumer CTA of the cluster
        }
        for (int a = 0; a < AS; ++a) {
            ptx::mbar_init(tfull + a, 1);
            ptx::mbar_init(tempty + a, 4);
        }
        ptx::fence_mbar_init();
    }
    __syncwarp();
    if (p.multicast) ptx::cluster_sync_all();  // peers signal these barriers: they must exist cluster-wide first
    if (warp == 4) {
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
        ptx::tc_fence_before_sync();
        __syncwarp();
        named_bar_arrive(1, kMmaThreads);
    } else {
        const int stid = tid < 128 ? tid : tid - 32;  // 0..159 over warps 0-3 and 5
        constexpr int NST = kMmaThreads - 32;
        grid_launch_dependents();  // lets the (PDL) candidate-reduce grid be scheduled as SMs drain
        if (!(NQ <= 32 && p.k <= 32)) {  // shared-memory lists exist only when the register path is not used
            list_init(L, NQ, stid, NST);
            // padding queries never produce candidates (same thread wrote tau[q] in list_init)
            for (int q = stid; q < NQ; q += NST)
                if (q >= nq) L.tau[q] = __int_as_float(0x7f800000);
        }
        // queries -> shared memory, K-major, 128B-swizzled, hi and lo parts.
        // unit of work: one 16-byte chunk (8 elements) of one query row.
        const int chunks_per_row = p.dim / 8;
        const int total = NQ * chunks_per_row;
        constexpr int QU = 4;  // chunks per thread whose (L2-latency) loads are issued together
        for (int idx0 = stid; idx0 < total; idx0 += NST * QU) {
            float4 v0[QU], v1[QU];
#pragma unroll
            for (int u = 0; u < QU; ++u) {
                const int idx = idx0 + u * NST;
                const int j = idx / chunks_per_row, cg = idx % chunks_per_row;
                v0[u] = v1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < total && j < nq) {
                    const float4 *src = reinterpret_cast<const float4 *>(qsrc + (long long)j * p.q_stride + cg * 8);
                    v0[u] = src[0];
                    v1[u] = src[1];
                }
            }
#pragma unroll
            for (int u = 0; u < QU; ++u) {
                const int idx = idx0 + u * NST;
                if (idx >= total) break;
                const int j = idx / chunks_per_row;   // query row
                const int cg = idx % chunks_per_row;  // global chunk
                const int kb = cg >> 3, c = cg & 7;
                uint4 hi, lo;
                const float x[8] = {v0[u].x, v0[u].y, v0[u].z, v0[u].w, v1[u].x, v1[u].y, v1[u].z, v1[u].w};
                split8<BF16>(x, p.lo_scale, hi, lo);  // zero rows (padding queries) stay zero
                unsigned char *tile = q_smem + (size_t)kb * NCOL * 128;
                *reinterpret_cast<uint4 *>(tile + j * 128 + ((c ^ (j & 7)) << 4)) = hi;
                if constexpr (SPLIT) {
                    const int jl = NQ + j;
                    *reinterpret_cast<uint4 *>(tile + jl * 128 + ((c ^ (jl & 7)) << 4)) = lo;
                }
            }
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        ptx::tc_fence_before_sync();
        __syncwarp();  // the staging loops above leave the lanes diverged
        named_bar_sync(1, kMmaThreads);
        ptx::tc_fence_after_sync();
        // broadcast through a shuffle so the compiler keeps the TMEM base in a uniform register
        tmem_base = __shfl_sync(kFullMask, *reinterpret_cast<volatile uint32_t *>(tmem_slot), 0);
    }

    // ---- roles: whole warps run the role loops (uniform control flow and uniform operands
    // for UTMALDG / UTCHMMA); one elected lane issues the asynchronous instructions ----------
    if (warp == 4) {
        uint32_t it = 0;
        const bool dyn = p.tile_ctr != nullptr;
        int tile = stream0 < p.n_tiles ? stream0 : -1;
        if (dyn && lane == 0) atomicMax(p.tile_ctr, (unsigned long long)p.epoch << 32);  // stale / zero -> (epoch, 0)
        for (uint32_t lt = 0;; ++lt) {
            unsigned long long ticket = 0;
            if (dyn && lane == 0) {
                tile_q[lt % kTileRing] = tile;
                __threadfence_block();
                *tile_pub = lt + 1;
                // ticket of tile lt + 1, taken now and looked at after this tile's loads are issued
                if (tile >= 0) ticket = atomicAdd(p.tile_ctr, 1ull);
            }
            if (tile < 0) break;
            for (int kg = 0; kg < KG; ++kg, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                ptx::mbar_wait(empty + s, ph ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(full + s, stage_bytes);
                    if (p.multicast) {
                        for (int j = 0; j < KPS; ++j)
                            ptx::tma_load_2d_multicast(
                                a_smem + (size_t)s * stage_bytes + (size_t)j * kStageBytes + (size_t)grp * slice_rows * 128,
                                &tmap_docs, (kg * KPS + j) * kBlockK, tile * kTileRows + grp * slice_rows, full + s,
                                cta_mask, p.tma_policy);
                    } else {
                        for (int j = 0; j < KPS; ++j)
                            ptx::tma_load_2d(a_smem + (size_t)s * stage_bytes + (size_t)j * kStageBytes, &tmap_docs,
                                             (kg * KPS + j) * kBlockK, tile * kTileRows, full + s, p.tma_policy);
                    }
                }
                __syncwarp();
                if (it == 0) timeline_stamp(p.timeline, 1);
            }
            if (dyn) {
                if (lane == 0) tile = tile_from_ticket(p.tile_ctr, ticket, p.epoch, p.n_tiles);
                tile = __shfl_sync(kFullMask, tile, 0);
            } else {
                const long long t = stream0 + (long long)(lt + 1) * n_streams;
                tile = t < p.n_tiles ? (int)t : -1;
            }
        }
        timeline_stamp(p.timeline, 2);
    } else if (warp == 5) {
        uint32_t it = 0;
        uint32_t lt = 0;  // local tile counter
        const uint32_t q_base = ptx::smem_u32(q_smem);
        const uint32_t a_base = ptx::smem_u32(a_smem);
        timeline_stamp(p.timeline, 3);
        for (;; ++lt) {
            if (consumer_tile(p, tile_pub, tile_q, lt, stream0, n_streams) < 0) break;
            const int as = lt % AS;
            const uint32_t aph = (lt / AS) & 1;
            ptx::mbar_wait(tempty + as, aph ^ 1);
            ptx::tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * NCOL);
            for (int kg = 0; kg < KG; ++kg, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                ptx::mbar_wait(full + s, ph);
                ptx::tc_fence_after_sync();
                if (ptx::elect_one()) {
                    for (int j = 0; j < KPS; ++j) {
                        const int kb = kg * KPS + j;
                        const uint64_t da0 = ptx::umma_desc_k_sw128(a_base + (uint32_t)s * stage_bytes + (uint32_t)j * kStageBytes);
                        const uint64_t db0 = ptx::umma_desc_k_sw128(q_base + (uint32_t)kb * (NCOL * 128));
#pragma unroll
                        for (int k4 = 0; k4 < kBlockK / 16; ++k4)  // +32 bytes along K = +2 in the address field
                            ptx::umma_f16(d_tmem, da0 + (uint64_t)(k4 * 2), db0 + (uint64_t)(k4 * 2), IDESC,
                                          (kb | k4) != 0 ? 1u : 0u);
                    }
                    // smem stage reusable once these MMAs retire -- in every CTA that multicasts into it
                    if (p.multicast) ptx::umma_commit_multicast(empty + s, cta_mask);
                    else ptx::umma_commit(empty + s);
                    if (kg == KG - 1) ptx::umma_commit(tfull + as);  // accumulator ready for the epilogue
                }
                __syncwarp();
            }
        }
        timeline_stamp(p.timeline, 4);
    } else if (NQ <= 32 && p.k <= 32) {
        // epilogue warps 0..3 (TMEM lane quadrant = warp), per-warp top-k lists in REGISTERS
        constexpr int RQ = NQ <= 32 ? NQ : 1;  // (keeps the arrays small when this branch is dead)
        float ls[RQ], tau[RQ];
        uint32_t li[RQ];
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            ls[q] = neg_inf();
            li[q] = invalid_id<uint32_t>();
            tau[q] = q < nq ? neg_inf() : __int_as_float(0x7f800000);
        }
        uint32_t lt = 0;
        if (warp == 0) timeline_stamp(p.timeline, 5);
        // one tile's scores against the lists: per 16-column group one compare per query and one ballot in the common
        // case; a passing row is inserted with one ballot + shuffle, or by a 32-wide bitonic merge when >= 4 lanes pass
        auto process = [&](int tile, int as) {
            const long long row = (long long)tile * kTileRows + warp * 32 + lane;
            const bool valid = row < p.n_rows;
            const uint32_t base_row = (uint32_t)(tile * kTileRows + warp * 32);
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * NCOL);
#pragma unroll
            for (int c0 = 0; c0 < RQ; c0 += 16) {
                float v[16];
                load_scores16<NCOL, SPLIT>(taddr, c0, p.lo_inv_scale, v);
                bool any = false;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < RQ) any |= (v[j] >= tau[c0 + j]);
                if (__ballot_sync(kFullMask, any && valid) == 0) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (c0 + j < RQ) {
                        const int q = c0 + j;
                        const bool pass = valid && v[j] >= tau[q];
                        unsigned m = __ballot_sync(kFullMask, pass);
                        if (m != 0) {
                            if (__popc(m) >= 4) {
                                const Entry e = reglist_merge32(ls[q], li[q], pass ? v[j] : neg_inf(),
                                                                pass ? base_row + lane : invalid_id<uint32_t>(), false);
                                ls[q] = e.s;
                                li[q] = e.i;
                            } else {
                                while (m) {
                                    const int src = __ffs(m) - 1;
                                    m &= m - 1;
                                    const Entry e = reglist_insert_one(ls[q], li[q], __shfl_sync(kFullMask, v[j], src),
                                                                       base_row + src);
                                    ls[q] = e.s;
                                    li[q] = e.i;
                                }
                            }
                            const uint32_t last = __shfl_sync(kFullMask, li[q], p.k - 1);
                            const float ts = __shfl_sync(kFullMask, ls[q], p.k - 1);
                            if (last != invalid_id<uint32_t>() && ts > tau[q]) {
                                tau[q] = ts;
                                if (tau_g != nullptr && lane == 0) atomicMax(tau_g + q, tau_encode(ts, p.epoch));
                            }
                        }
                    }
                }
            }
        };
        auto release = [&](int as) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty + as);
        };
        // warm-up seed (see MmaParams::slot_g): the first tile publishes per-query maxima and keeps its stage
        const bool seed = p.slot_g != nullptr && tau_g != nullptr;
        unsigned long long *slots = seed ? p.slot_g + (long long)q0 * kSlotStride : nullptr;
        const int my_slot = (stream0 * 4 + warp) % p.k;
        int held_tile = -1;
        for (;; ++lt) {
            const int tile = consumer_tile(p, tile_pub, tile_q, lt, stream0, n_streams);
            if (tile < 0) break;
            const int as = lt % AS;
            const uint32_t aph = (lt / AS) & 1;
            // pick up the other CTAs' thresholds (one coalesced load per warp, issued before the wait)
            unsigned long long graw = 0;
            if (tau_g != nullptr && lane < nq) graw = ld_volatile_u64(tau_g + lane);
            ptx::mbar_wait(tfull + as, aph);
            ptx::tc_fence_after_sync();
            if (tau_g != nullptr) {
                const float tg = tau_decode(graw, p.epoch);
#pragma unroll
                for (int q = 0; q < RQ; ++q) tau[q] = fmaxf(tau[q], __shfl_sync(kFullMask, tg, q));
            }
            if (seed && lt == 0) {
                const long long row = (long long)tile * kTileRows + warp * 32 + lane;
                const bool valid = row < p.n_rows;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * NCOL);
#pragma unroll
                for (int c0 = 0; c0 < RQ; c0 += 16) {
                    float v[16];
                    load_scores16<NCOL, SPLIT>(taddr, c0, p.lo_inv_scale, v);
                    float mine = neg_inf();   // lane j ends up with the warp maximum of query c0 + j
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float mx = valid ? v[j] : neg_inf();
#pragma unroll
                        for (int sh = 16; sh >= 1; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, sh));
                        if (lane == j) mine = mx;
                    }
                    if (lane < 16 && c0 + lane < nq && mine > neg_inf())
                        atomicMax(slots + (long long)(c0 + lane) * kSlotStride + my_slot, tau_encode(mine, p.epoch));
                }
                held_tile = tile;   // stage 0 is released after the bounds have been read (next tile, or the end)
                continue;
            }
            if (held_tile >= 0) {
                // the k slots of each query hold the best scores of k disjoint sets of rows: their minimum is a lower
                // bound of the global k-th best (an empty / stale slot decodes to -inf: no bound yet)
                // (all loads first, sixteen queries at a time: a volatile load may not pass the atomic that follows the
                // previous query's minimum, and 32 dependent L2 round trips cost ~13 us -- profiles/r2_call4.log)
#pragma unroll
                for (int h0 = 0; h0 < RQ; h0 += 16) {
                    unsigned long long raw[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        raw[j] = 0ull;
                        if (h0 + j < RQ && h0 + j < nq && lane < p.k)
                            raw[j] = ld_volatile_u64(slots + (long long)(h0 + j) * kSlotStride + lane);
                    }
                    float bound[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float b = lane < p.k ? tau_decode(raw[j], p.epoch) : __int_as_float(0x7f800000);
#pragma unroll
                        for (int sh = 16; sh >= 1; sh >>= 1) b = fminf(b, __shfl_xor_sync(kFullMask, b, sh));
                        bound[j] = b;
                    }
                    float mine = neg_inf();   // lane j publishes query h0 + j's bound if it improves the threshold
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (h0 + j < RQ) {
                            const int q = h0 + j;
                            if (q < nq && bound[j] > tau[q]) {
                                tau[q] = bound[j];
                                if (lane == j) mine = bound[j];
                            }
                        }
                    }
                    if (lane < 16 && mine > neg_inf()) atomicMax(tau_g + h0 + lane, tau_encode(mine, p.epoch));
                }
                process(held_tile, 0);
                release(0);
                held_tile = -1;
            }
            process(tile, as);
            release(as);
            if (p.timeline != nullptr && warp == 0 && ((lt + 1) & lt) == 0 && lt < 64)  // tiles 0, 1, 3, 7, 15, 31, 63
                timeline_stamp(p.timeline, 6 + (31 - __clz((int)lt + 1)));
        }
        if (held_tile >= 0) {  // a CTA with a single tile: no second tile came to trigger the deferred pass
            process(held_tile, 0);
            release(0);
        }
        if (warp == 0) timeline_stamp(p.timeline, 13);
        // Every MMA (hence every TMA write) of this CTA has retired once the last accumulator was
        // handed over: the document ring is free, park this warp's lists there for the CTA merge.
        float *ms = reinterpret_cast<float *>(a_smem);
        uint32_t *mi = reinterpret_cast<uint32_t *>(a_smem + 4 * RQ * 32 * sizeof(float));
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            ms[(warp * RQ + q) * 32 + lane] = ls[q];
            mi[(warp * RQ + q) * 32 + lane] = li[q];
        }
    } else {
        // epilogue warps 0..3: TMEM lane quadrant = warp; CTA-shared lists in shared memory (k > 32)
        uint32_t lt = 0;
        for (;; ++lt) {
            const int tile = consumer_tile(p, tile_pub, tile_q, lt, stream0, n_streams);
            if (tile < 0) break;
            const int as = lt % AS;
            const uint32_t aph = (lt / AS) & 1;
            ptx::mbar_wait(tfull + as, aph);
            ptx::tc_fence_after_sync();
            const long long row = (long long)tile * kTileRows + warp * 32 + lane;
            const bool valid = row < p.n_rows;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * NCOL);
#pragma unroll 1
            for (int c0 = 0; c0 < NQ; c0 += 16) {
                float v[16];
                load_scores16<NCOL, SPLIT>(taddr, c0, p.lo_inv_scale, v);
                bool any = false;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (c0 + j < NQ) {
                        const float thr = *(volatile float *)(L.tau + c0 + j);
                        any |= (v[j] >= thr);
                    }
                }
                any = any && valid;
                if (__ballot_sync(kFullMask, any) == 0) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int qi = c0 + j;
                    if (qi >= nq) break;  // warp-uniform
                    const float thr = *(volatile float *)(L.tau + qi);
                    const bool pass = valid && v[j] >= thr;
                    unsigned m = __ballot_sync(kFullMask, pass);
                    if (m == 0) continue;
                    const uint32_t base_row = (uint32_t)(tile * kTileRows + warp * 32);
                    if (L.kpl == 1 && __popc(m) >= 6) {
                        list_insert_bulk32(L, qi, pass ? v[j] : neg_inf(),
                                           pass ? base_row + lane : invalid_id<uint32_t>());
                    } else {
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            list_insert<uint32_t>(L, qi, __shfl_sync(kFullMask, v[j], src), base_row + src);
                        }
                    }
                }
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty + as);
        }
    }

    // ---- teardown --------------------------------------------------------------
    ptx::tc_fence_before_sync();
    __syncthreads();
    float *cs = p.cand_s + (long long)blockIdx.x * p.cand_stride + (long long)q0 * p.k;
    uint32_t *ci = p.cand_i + (long long)blockIdx.x * p.cand_stride + (long long)q0 * p.k;
    if (NQ <= 32 && p.k <= 32) {
        // merge the four epilogue warps' lists per query (bitonic, in registers) and publish
        const float *ms = reinterpret_cast<const float *>(a_smem);
        const uint32_t *mi = reinterpret_cast<const uint32_t *>(a_smem + 4 * NQ * 32 * sizeof(float));
        // every warp takes the queries q = warp, warp + 6, ... and merges them TOGETHER (interleaved chains, see
        // reglist_merge32_batch): the four warps' lists arrive sorted, so each merge is 5 compare-exchange steps
        constexpr int NW = kMmaThreads / 32;
        constexpr int TQ = (NQ + NW - 1) / NW;
        float s0[TQ];
        uint32_t i0[TQ];
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
            const int q = warp + NW * t < NQ ? warp + NW * t : 0;
            s0[t] = ms[q * 32 + lane];
            i0[t] = mi[q * 32 + lane];
        }
        for (int w = 1; w < 4; ++w) {
            float c_s[TQ];
            uint32_t c_i[TQ];
#pragma unroll
            for (int t = 0; t < TQ; ++t) {
                const int q = warp + NW * t < NQ ? warp + NW * t : 0;
                c_s[t] = ms[(w * NQ + q) * 32 + lane];
                c_i[t] = mi[(w * NQ + q) * 32 + lane];
            }
            reglist_merge32_batch<TQ, TQ>(s0, i0, 0, c_s, c_i, true);
        }
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
            const int q = warp + NW * t;
            if (q < nq && lane < p.k) {
                cs[q * p.k + lane] = s0[t];
                ci[q * p.k + lane] = i0[t];
            }
        }
    } else {
        for (int idx = tid; idx < nq * p.k; idx += kMmaThreads) {
            const int b = idx / p.k, e = idx % p.k;
            cs[idx] = L.s[b * L.kcap + e];
            ci[idx] = L.i[b * L.kcap + e];
        }
    }
    if (warp == 0) timeline_stamp(p.timeline, 14);
    if (warp == 4) {
        __syncwarp();
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(*reinterpret_cast<volatile uint32_t *>(tmem_slot), TMEM_COLS);
    }
    // peers may still multicast-commit into this CTA's barriers: nobody leaves before everybody is done
    if (p.multicast) ptx::cluster_sync_all();
}

}  // namespace vqa
