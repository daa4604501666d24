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
// on the fly from fp32 to the storage type as a hi + lo pair; both parts are
// multiplied against the same document block and accumulate into the SAME TMEM
// columns, so that the only rounding left is the documents' own storage
// rounding -- this is what keeps recall@k >= 0.999 against the fp32 verify mode.
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
    int nq;  // queries in this pass: nq <= NCOL
    int k;
    long long n_rows;
    int dim;  // multiple of 64
    int split;  // 1: hi + lo query parts (2 MMAs per K step), 0: hi only
    float *cand_s;
    uint32_t *cand_i;
    long long cand_stride;
    int n_tiles;
    int n_stages;  // smem ring depth
    int kps;       // k-blocks (16 KB TMA boxes) per ring stage
    unsigned long long tma_policy;  // L2 cache hint for the document stream
};

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
__device__ __forceinline__ void split8(const float (&x)[8], uint4 &hi, uint4 &lo) {
    float h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if constexpr (BF16) {
            h[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
        } else {
            h[j] = __half2float(__float2half_rn(x[j]));
        }
        l[j] = x[j] - h[j];  // exact in fp32; fp16's residual may be subnormal (still ~2^-19 relative overall)
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

__device__ __noinline__ void list_insert_bulk32(ListView<uint32_t> L, int q, float cs, uint32_t ci) {
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

template <bool BF16, int NCOL>
__global__ void __launch_bounds__(kMmaThreads, 1)
mma_topk_kernel(const __grid_constant__ CUtensorMap tmap_docs, const MmaParams p) {
    constexpr int AS = mma_acc_stages<NCOL>();
    constexpr uint32_t TMEM_COLS = AS * NCOL;  // power of two, 128..512
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
    unsigned char *q_smem = smem;                                   // KB x {hi, lo} tiles of NCOL*128 B
    unsigned char *a_smem = q_smem + (size_t)KB * 2 * NCOL * 128;   // S stages of KPS x 16 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(a_smem + (size_t)S * stage_bytes);
    uint64_t *full = bars;                       // [kMaxStages]
    uint64_t *empty = bars + kMaxStages;         // [kMaxStages]
    uint64_t *tfull = bars + 2 * kMaxStages;     // [kMaxAccStages]
    uint64_t *tempty = tfull + kMaxAccStages;    // [kMaxAccStages]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + kMaxAccStages);
    ListView<uint32_t> L = list_carve<uint32_t>(reinterpret_cast<unsigned char *>(bars) + 1024, NCOL, p.k);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    // ---- one-time setup ------------------------------------------------------
    if (warp == 5 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(full + s, 1);
            ptx::mbar_init(empty + s, 1);
        }
        for (int a = 0; a < AS; ++a) {
            ptx::mbar_init(tfull + a, 1);
            ptx::mbar_init(tempty + a, 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 4) {
        if (lane == 0) ptx::prefetch_tmap(&tmap_docs);
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    list_init(L, NCOL, tid, kMmaThreads);
    grid_launch_dependents();  // lets the (PDL) candidate-reduce grid be scheduled as SMs drain
    // padding columns never produce candidates (same thread wrote tau[q] in list_init)
    for (int q = tid; q < NCOL; q += kMmaThreads)
        if (q >= p.nq) L.tau[q] = __int_as_float(0x7f800000);

    // queries -> shared memory, K-major, 128B-swizzled, hi (and lo) parts.
    // unit of work: one 16-byte chunk (8 elements) of one query row.
    {
        const int chunks_per_row = p.dim / 8;
        const int total = NCOL * chunks_per_row;
        for (int idx = tid; idx < total; idx += kMmaThreads) {
            const int j = idx / chunks_per_row;   // query row
            const int cg = idx % chunks_per_row;  // global chunk
            const int kb = cg >> 3, c = cg & 7;
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (j < p.nq) {
                const float4 *src = reinterpret_cast<const float4 *>(p.q + (long long)j * p.q_stride + cg * 8);
                const float4 v0 = src[0], v1 = src[1];
                const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                split8<BF16>(x, hi, lo);
            }
            unsigned char *tile = q_smem + (size_t)kb * 2 * NCOL * 128;
            const int off = j * 128 + ((c ^ (j & 7)) << 4);
            *reinterpret_cast<uint4 *>(tile + off) = hi;
            *reinterpret_cast<uint4 *>(tile + NCOL * 128 + off) = lo;
        }
    }
    ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    // ---- roles ---------------------------------------------------------------
    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int kg = 0; kg < KG; ++kg, ++it) {
                    const int s = it % S;
                    const uint32_t ph = (it / S) & 1;
                    ptx::mbar_wait(empty + s, ph ^ 1);
                    ptx::mbar_arrive_expect_tx(full + s, stage_bytes);
                    for (int j = 0; j < KPS; ++j)
                        ptx::tma_load_2d(a_smem + (size_t)s * stage_bytes + (size_t)j * kStageBytes, &tmap_docs,
                                         (kg * KPS + j) * kBlockK, tile * kTileRows, full + s, p.tma_policy);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (lane == 0) {
            uint32_t it = 0;
            uint32_t lt = 0;  // local tile counter
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
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
                    for (int j = 0; j < KPS; ++j) {
                        const int kb = kg * KPS + j;
                        const uint32_t a_addr = ptx::smem_u32(a_smem + (size_t)s * stage_bytes + (size_t)j * kStageBytes);
                        const uint32_t b_addr = ptx::smem_u32(q_smem + (size_t)kb * 2 * NCOL * 128);
#pragma unroll
                        for (int k4 = 0; k4 < kBlockK / 16; ++k4) {
                            const uint64_t da = ptx::umma_desc_k_sw128(a_addr + k4 * 32);
                            const uint64_t db = ptx::umma_desc_k_sw128(b_addr + k4 * 32);
                            ptx::umma_f16(d_tmem, da, db, IDESC, (kb | k4) != 0 ? 1u : 0u);
                            if (p.split) {  // + docs x q_lo into the same accumulator columns
                                const uint64_t dl = ptx::umma_desc_k_sw128(b_addr + NCOL * 128 + k4 * 32);
                                ptx::umma_f16(d_tmem, da, dl, IDESC, 1u);
                            }
                        }
                    }
                    ptx::umma_commit(empty + s);  // smem stage reusable once these MMAs retire
                }
                ptx::umma_commit(tfull + as);  // accumulator ready for the epilogue
            }
        }
        __syncwarp();
    } else {
        // epilogue warps 0..3: TMEM lane quadrant = warp
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
            const int as = lt % AS;
            const uint32_t aph = (lt / AS) & 1;
            ptx::mbar_wait(tfull + as, aph);
            ptx::tc_fence_after_sync();
            const long long row = (long long)tile * kTileRows + warp * 32 + lane;
            const bool valid = row < p.n_rows;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * NCOL);
            for (int c0 = 0; c0 < NCOL; c0 += 16) {
                uint32_t acc[16];
                ptx::tmem_ld16(taddr + c0, acc);
                ptx::tmem_ld_wait();
                float v[16];
                bool any = false;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = __uint_as_float(acc[j]);
                    const float thr = *(volatile float *)(L.tau + c0 + j);
                    any |= (v[j] >= thr);
                }
                any = any && valid;
                if (__ballot_sync(kFullMask, any) == 0) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int qi = c0 + j;
                    if (qi >= p.nq) break;  // warp-uniform
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
    float *cs = p.cand_s + (long long)blockIdx.x * p.cand_stride;
    uint32_t *ci = p.cand_i + (long long)blockIdx.x * p.cand_stride;
    for (int idx = tid; idx < p.nq * p.k; idx += kMmaThreads) {
        const int b = idx / p.k, e = idx % p.k;
        cs[idx] = L.s[b * L.kcap + e];
        ci[idx] = L.i[b * L.kcap + e];
    }
    if (warp == 4) {
        __syncwarp();
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace vqa
