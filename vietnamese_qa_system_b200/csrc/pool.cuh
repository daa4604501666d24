// pool.cuh -- K1: fused masked mean-pool + L2-normalise; stand-alone row
// normalise (+ cast to the index storage type); the two-index agreement rule.
//
// Replaces txtai MeanPooling.forward + normalize (numpy) under Embeddings.index /
// Embeddings.search (heavy_ranker.py:86,88,98,100; in-tree twin src/test.py:97-99):
//     e[b,:] = sum_s h[b,s,:]*m[b,s] / max(sum_s m[b,s], 1e-9);   e /= ||e||_2
// One pass over the hidden states (HBM roofline: B*S_valid*D*sizeof(h) bytes;
// fully masked tokens are never loaded), fp32 accumulation.
#pragma once

#include "common.cuh"

namespace vqa {

constexpr int kPoolThreads = 256;

template <typename MT>
__device__ __forceinline__ float mask_to_float(MT v) { return (float)v; }

__device__ __forceinline__ float block_sum(float v, float *red /*[32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(kFullMask, v, m);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    if (warp == 0) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) t += __shfl_xor_sync(kFullMask, t, m);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    t = red[0];
    return t;
}

// grid = B.  dynamic smem: G*dim floats (partials) + dim floats (pooled) + 32 floats.
template <typename T, typename MT>
__global__ void __launch_bounds__(kPoolThreads)
pool_normalize_kernel(const unsigned char *__restrict__ hidden, const MT *__restrict__ mask, int seq, int dim,
                      int normalize, float *__restrict__ out) {
    constexpr int E = Elem<T>::E;
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int nchunks = dim / E;
    const int cw = nchunks < kPoolThreads ? nchunks : kPoolThreads;  // chunk lanes
    const int G = kPoolThreads / cw;                                  // token groups
    float *part = reinterpret_cast<float *>(smem);                    // [G][dim]
    float *pooled = part + (size_t)G * dim;                           // [dim]
    float *red = pooled + dim;                                        // [32]

    const MT *mrow = mask + (long long)b * seq;
    const unsigned char *hrow = hidden + (long long)b * seq * dim * sizeof(T);

    // token count (sum of mask weights)
    float cnt = 0.f;
    for (int s = tid; s < seq; s += kPoolThreads) cnt += mask_to_float(mrow[s]);
    cnt = block_sum(cnt, red);

    const int g = tid / cw;
    const int cl = tid % cw;
    if (g < G) {
        for (int c = cl; c < nchunks; c += cw) {
            float acc[E];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = 0.f;
            int s = g;
            for (; s + 3 * G < seq; s += 4 * G) {
                float m[4];
                uint4 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) m[u] = mask_to_float(mrow[s + u * G]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    w[u] = make_uint4(0, 0, 0, 0);
                    if (m[u] != 0.f)
                        w[u] = ldg_stream(hrow + ((long long)(s + u * G) * dim + (long long)c * E) * sizeof(T));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float x[E];
                    Elem<T>::unpack(w[u], x);
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(x[e], m[u], acc[e]);
                }
            }
            for (; s < seq; s += G) {
                const float m = mask_to_float(mrow[s]);
                if (m != 0.f) {
                    const uint4 w = ldg_stream(hrow + ((long long)s * dim + (long long)c * E) * sizeof(T));
                    float x[E];
                    Elem<T>::unpack(w, x);
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(x[e], m, acc[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e) part[(size_t)g * dim + c * E + e] = acc[e];
        }
    }
    __syncthreads();
    const float den = fmaxf(cnt, 1e-9f);
    float ss = 0.f;
    for (int d = tid; d < dim; d += kPoolThreads) {
        float t = 0.f;
        for (int gg = 0; gg < G; ++gg) t += part[(size_t)gg * dim + d];
        const float mean = t / den;
        pooled[d] = mean;
        ss += mean * mean;
    }
    ss = block_sum(ss, red);
    float *orow = out + (long long)b * dim;
    if (normalize) {
        const float nrm = sqrtf(ss);
        for (int d = tid; d < dim; d += kPoolThreads) orow[d] = nrm > 0.f ? pooled[d] / nrm : 0.f;
    } else {
        for (int d = tid; d < dim; d += kPoolThreads) orow[d] = pooled[d];
    }
}

// mask element by runtime dtype code (vqa_dtype): I64=3, I32=4, U8=5, F32=0
__device__ __forceinline__ float mask_at(const void *m, int mdt, long long idx) {
    switch (mdt) {
        case 3: return (float)static_cast<const long long *>(m)[idx];
        case 4: return (float)static_cast<const int *>(m)[idx];
        case 5: return (float)static_cast<const unsigned char *>(m)[idx];
        default: return static_cast<const float *>(m)[idx];
    }
}

// Fast path (row of <= ITERS*512 bytes): one warp per TOKEN, lanes stride over the row's 16-byte
// chunks, up to 4 tokens in flight per warp (as many as the register budget holds: ITERS * kPoolUnroll
// independent 128-bit loads per lane, 6 at dim 768 bf16), 16 warps per CTA
// splitting the sequence, fp32 partial sums combined through shared memory.  grid = B.
// dynamic smem: (16 + 1) * dim + 32 + 16 floats.
constexpr int kPoolFastThreads = 512;
constexpr int kPoolWarps = kPoolFastThreads / 32;

template <typename T, int ITERS>
__global__ void __launch_bounds__(kPoolFastThreads, (ITERS * (16 / (int)sizeof(T)) <= 24) ? 2 : 1)
pool_normalize_warp_kernel(const unsigned char *__restrict__ hidden, const void *__restrict__ mask, int mdt, int seq,
                           int dim, int normalize, float *__restrict__ out) {
    constexpr int E = Elem<T>::E;
    // Tokens in flight per warp: the 16-byte loads (4 registers each) of kPoolUnroll tokens plus the ITERS * E
    // accumulators must fit the register budget (64 per thread at 2 CTAs per SM, 128 at 1) -- with a fixed 4
    // the 768-dim bf16 instantiation spilled ~400 bytes per thread inside the token loop.
    constexpr int kRegBudget = (ITERS * E <= 24) ? 44 : 100;
    constexpr int kFit = (kRegBudget - ITERS * E) / (ITERS * 4);
    constexpr int kPoolUnroll = kFit < 1 ? 1 : (kFit > 4 ? 4 : kFit);
    extern __shared__ __align__(16) unsigned char smem[];
    float *part = reinterpret_cast<float *>(smem);          // [kPoolWarps][dim]
    float *pooled = part + (size_t)kPoolWarps * dim;        // [dim]
    float *red = pooled + dim;                              // [32]
    float *cntw = red + 32;                                 // [kPoolWarps]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = dim / E;
    const unsigned char *hrow = hidden + (long long)b * seq * dim * sizeof(T);
    const long long mbase = (long long)b * seq;

    float acc[ITERS][E];
#pragma unroll
    for (int i = 0; i < ITERS; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[i][e] = 0.f;
    float cnt = 0.f;
    // Tokens after the last unmasked one (right padding, the usual layout) are not visited at all:
    // s_end = 1 + last index with a non-zero mask weight.
    int last = -1;
    for (int s = tid; s < seq; s += kPoolFastThreads)
        if (mask_at(mask, mdt, mbase + s) != 0.f) last = s;
    last = __reduce_max_sync(kFullMask, last);
    int *lastw = reinterpret_cast<int *>(cntw);  // (cntw is written only after the token loop)
    if (lane == 0) lastw[warp] = last;
    __syncthreads();
    int s_end = 0;
#pragma unroll
    for (int w2 = 0; w2 < kPoolWarps; ++w2) s_end = max(s_end, lastw[w2] + 1);
    __syncthreads();
    for (int s0 = warp * kPoolUnroll; s0 < s_end; s0 += kPoolWarps * kPoolUnroll) {
        float m[kPoolUnroll];
        uint4 w[kPoolUnroll][ITERS];
#pragma unroll
        for (int u = 0; u < kPoolUnroll; ++u) m[u] = s0 + u < s_end ? mask_at(mask, mdt, mbase + s0 + u) : 0.f;
#pragma unroll
        for (int u = 0; u < kPoolUnroll; ++u) {
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                w[u][i] = make_uint4(0, 0, 0, 0);
                const int c = i * 32 + lane;
                if (m[u] != 0.f && c < nchunks)
                    w[u][i] = ldg_stream(hrow + ((long long)(s0 + u) * dim + (long long)c * E) * sizeof(T));
            }
        }
#pragma unroll
        for (int u = 0; u < kPoolUnroll; ++u) {
            cnt += m[u];
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                float x[E];
                Elem<T>::unpack(w[u][i], x);
#pragma unroll
                for (int e = 0; e < E; ++e) acc[i][e] = fmaf(x[e], m[u], acc[i][e]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
        const int c = i * 32 + lane;
        if (c < nchunks) {
#pragma unroll
            for (int e = 0; e < E; ++e) part[(size_t)warp * dim + c * E + e] = acc[i][e];
        }
    }
    if (lane == 0) cntw[warp] = cnt;
    __syncthreads();
    float total_cnt = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kPoolWarps; ++w2) total_cnt += cntw[w2];
    const float den = fmaxf(total_cnt, 1e-9f);
    float ss = 0.f;
    for (int d = tid; d < dim; d += kPoolFastThreads) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kPoolWarps; ++w2) t += part[(size_t)w2 * dim + d];
        const float mean = t / den;
        pooled[d] = mean;
        ss += mean * mean;
    }
    ss = block_sum(ss, red);
    float *orow = out + (long long)b * dim;
    if (normalize) {
        const float nrm = sqrtf(ss);
        for (int d = tid; d < dim; d += kPoolFastThreads) orow[d] = nrm > 0.f ? pooled[d] / nrm : 0.f;
    } else {
        for (int d = tid; d < dim; d += kPoolFastThreads) orow[d] = pooled[d];
    }
}

// Row-wise L2 normalise: one warp per row, float4 accesses.  Optional cast copy.
// cast_kind: 0 none, 1 bf16, 2 f16.
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float *__restrict__ in, long long in_stride, long long n_rows, int dim,
                      float *__restrict__ out, long long out_stride, void *__restrict__ cast_out, int cast_kind,
                      long long cast_stride) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float4 *src = reinterpret_cast<const float4 *>(in + row * in_stride);
    const int n4 = dim >> 2;
    float ss = 0.f;
    for (int c = lane; c < n4; c += 32) {
        const float4 v = src[c];
        ss = fmaf(v.x, v.x, ss);
        ss = fmaf(v.y, v.y, ss);
        ss = fmaf(v.z, v.z, ss);
        ss = fmaf(v.w, v.w, ss);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(kFullMask, ss, m);
    const float nrm = sqrtf(ss);
    const bool ok = nrm > 0.f;
    for (int c = lane; c < n4; c += 32) {
        float4 v = src[c];
        v.x = ok ? v.x / nrm : 0.f;
        v.y = ok ? v.y / nrm : 0.f;
        v.z = ok ? v.z / nrm : 0.f;
        v.w = ok ? v.w / nrm : 0.f;
        if (out) reinterpret_cast<float4 *>(out + row * out_stride)[c] = v;
        if (cast_kind == 1) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b2 = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b2));
            reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(cast_out) + row * cast_stride)[c] = pk;
        } else if (cast_kind == 2) {
            __half2 a = __floats2half2_rn(v.x, v.y), b2 = __floats2half2_rn(v.z, v.w);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b2));
            reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(cast_out) + row * cast_stride)[c] = pk;
        }
    }
}

// heavy_ranker.py:110 -- accept iff same id and score_a + score_b > threshold
// (Python adds two floats in double precision).
__global__ void agree_kernel(const long long *ids_a, const float *sa, const long long *ids_b, const float *sb,
                             long long n, double threshold, unsigned char *accept, float *combined) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sum = (double)sa[i] + (double)sb[i];
    if (accept) accept[i] = (ids_a[i] == ids_b[i] && sum > threshold) ? 1 : 0;
    if (combined) combined[i] = (float)sum;
}

}  // namespace vqa
