// common.cuh -- shared device helpers: ordering, CTA-shared top-k lists, loads.
//
// Ordering contract (include/vqa.h): score descending, ties -> lower id.  Every
// selection level (per-CTA list, cross-CTA reduce, cross-rank merge) uses the
// same comparator, so the result is independent of how rows are partitioned.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "consts.h"

namespace vqa {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// programmatic dependent launch (PDL): the consumer grid blocks in grid_dependency_wait() until the
// producer grid has finished and its writes are visible; the producer may release the launch early.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename IdT>
__device__ __forceinline__ constexpr IdT invalid_id();
template <>
__device__ __forceinline__ constexpr uint32_t invalid_id<uint32_t>() { return 0xffffffffu; }
template <>
__device__ __forceinline__ constexpr long long invalid_id<long long>() { return 0x7fffffffffffffffLL; }

// a ranks strictly before b
template <typename IdT>
__device__ __forceinline__ bool ranks_before(float as, IdT ai, float bs, IdT bi) {
    return (as > bs) || (as == bs && ai < bi);
}

// ---------------------------------------------------------------------------
// CTA-shared sorted top-k lists in shared memory, one per query, guarded by a
// per-query spin lock taken by a whole warp.  Insertions are rare (a candidate
// must first beat the published threshold tau[q]), so contention is negligible.
// Layout: s[q][kcap], i[q][kcap] with kcap = 32*ceil(k/32); entries >= k and
// unfilled entries hold (-inf, invalid) and are never overwritten past k.
// ---------------------------------------------------------------------------
template <typename IdT>
struct ListView {
    float *s;
    IdT *i;
    float *tau;  // tau[q] = score of entry k-1 (or -inf while the list is not full)
    int *lock;
    int k;
    int kcap;
    int kpl;  // kcap / 32
};

template <typename IdT>
__host__ __device__ inline size_t list_smem_bytes(int nq, int k) {
    int kcap = ((k + 31) / 32) * 32;
    return (size_t)nq * kcap * (sizeof(float) + sizeof(IdT)) + (size_t)nq * (sizeof(float) + sizeof(int));
}

// carve from a 16-byte aligned smem base
template <typename IdT>
__device__ __forceinline__ ListView<IdT> list_carve(unsigned char *base, int nq, int k) {
    ListView<IdT> L;
    int kcap = ((k + 31) / 32) * 32;
    L.k = k;
    L.kcap = kcap;
    L.kpl = kcap / 32;
    L.i = reinterpret_cast<IdT *>(base);
    L.s = reinterpret_cast<float *>(base + (size_t)nq * kcap * sizeof(IdT));
    L.tau = L.s + (size_t)nq * kcap;
    L.lock = reinterpret_cast<int *>(L.tau + nq);
    return L;
}

// all `nthreads` threads (tid in [0,nthreads)) cooperate; caller syncs afterwards
template <typename IdT>
__device__ __forceinline__ void list_init(const ListView<IdT> &L, int nq, int tid, int nthreads) {
    for (int e = tid; e < nq * L.kcap; e += nthreads) {
        L.s[e] = neg_inf();
        L.i[e] = invalid_id<IdT>();
    }
    for (int q = tid; q < nq; q += nthreads) {
        L.tau[q] = neg_inf();
        L.lock[q] = 0;
    }
}

__device__ __forceinline__ long long shfl_any(long long v, int src) { return __shfl_sync(kFullMask, v, src); }
__device__ __forceinline__ uint32_t shfl_any(uint32_t v, int src) { return __shfl_sync(kFullMask, v, src); }
__device__ __forceinline__ long long shfl_up_any(long long v) { return __shfl_up_sync(kFullMask, v, 1); }
__device__ __forceinline__ uint32_t shfl_up_any(uint32_t v) { return __shfl_up_sync(kFullMask, v, 1); }

// Whole warp calls this with identical (q, cs, ci).  Inserts the candidate into
// list q if it ranks before the current k-th entry.
template <typename IdT>
__device__ __noinline__ void list_insert(ListView<IdT> L, int q, float cs, IdT ci) {
    const int lane = threadIdx.x & 31;
    volatile float *ls = L.s + (size_t)q * L.kcap;
    volatile IdT *li = L.i + (size_t)q * L.kcap;
    volatile float *tau = L.tau;
    if (lane == 0) {
        while (atomicCAS(L.lock + q, 0, 1) != 0) __nanosleep(32);
    }
    __syncwarp();
    __threadfence_block();

    float es[kMaxKpl];
    IdT ei[kMaxKpl];
    int p = 0;
#pragma unroll
    for (int j = 0; j < kMaxKpl; ++j) {
        es[j] = neg_inf();
        ei[j] = invalid_id<IdT>();
        if (j < L.kpl) {
            es[j] = ls[j * 32 + lane];
            ei[j] = li[j * 32 + lane];
            p += __popc(__ballot_sync(kFullMask, ranks_before<IdT>(es[j], ei[j], cs, ci)));
        }
    }
    if (p < L.k) {
#pragma unroll
        for (int j = kMaxKpl - 1; j >= 0; --j) {
            if (j < L.kpl) {
                float ups = __shfl_up_sync(kFullMask, es[j], 1);
                IdT upi = shfl_up_any(ei[j]);
                if (j > 0) {
                    float ws = __shfl_sync(kFullMask, es[j - 1], 31);
                    IdT wi = shfl_any(ei[j - 1], 31);
                    if (lane == 0) {
                        ups = ws;
                        upi = wi;
                    }
                }
                const int e = j * 32 + lane;
                float ns = es[j];
                IdT ni = ei[j];
                if (e == p) {
                    ns = cs;
                    ni = ci;
                } else if (e > p) {
                    ns = ups;
                    ni = upi;
                }
                if (e >= p && e < L.k) {
                    ls[e] = ns;
                    li[e] = ni;
                    if (e == L.k - 1) tau[q] = ns;
                }
            }
        }
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(L.lock + q, 0);
}

// ---------------------------------------------------------------------------
// streaming 128-bit load, read-only path, no L1 allocation
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 16-byte chunk -> fp32 elements (exact conversions)
template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static constexpr int E = 4;
    __device__ __forceinline__ static void unpack(const uint4 &w, float (&x)[4]) {
        x[0] = __uint_as_float(w.x);
        x[1] = __uint_as_float(w.y);
        x[2] = __uint_as_float(w.z);
        x[3] = __uint_as_float(w.w);
    }
};
template <>
struct Elem<__nv_bfloat16> {
    static constexpr int E = 8;
    __device__ __forceinline__ static void unpack(const uint4 &w, float (&x)[8]) {
        x[0] = __uint_as_float(w.x << 16);
        x[1] = __uint_as_float(w.x & 0xffff0000u);
        x[2] = __uint_as_float(w.y << 16);
        x[3] = __uint_as_float(w.y & 0xffff0000u);
        x[4] = __uint_as_float(w.z << 16);
        x[5] = __uint_as_float(w.z & 0xffff0000u);
        x[6] = __uint_as_float(w.w << 16);
        x[7] = __uint_as_float(w.w & 0xffff0000u);
    }
};
template <>
struct Elem<__half> {
    static constexpr int E = 8;
    __device__ __forceinline__ static void unpack(const uint4 &w, float (&x)[8]) {
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __half2 h = *reinterpret_cast<const __half2 *>(&ws[j]);
            float2 f = __half22float2(h);
            x[2 * j] = f.x;
            x[2 * j + 1] = f.y;
        }
    }
};

}  // namespace vqa
