// misc_launch.cu -- scan dispatch by dtype, cross-CTA / cross-rank reduce,
// pooling, normalise, agreement rule.
#include "../../include/vqa.h"
#include "launch.h"
#include "pool.cuh"
#include "scan.cuh"


namespace vqa {

constexpr size_t kSelectSmemLimit = 227 * 1024;  // dynamic shared memory one CTA may opt in to on sm_100

cudaError_t launch_scan(const ScanLaunch &a, cudaStream_t st) {
    switch (a.dtype) {
        case VQA_F32: return launch_scan_f32(a, st);
        case VQA_BF16: return launch_scan_bf16(a, st);
        case VQA_F16: return launch_scan_f16(a, st);
        default: return cudaErrorInvalidValue;
    }
}

template <typename IdT>
static cudaError_t launch_reduce_t(const float *cand_s, const IdT *cand_i, long long list_stride,
                                   long long list_stride_i, long long query_stride, int n_lists, int k_in, int k_out, long long id_base,
                                   float *out_s, long long *out_i, int n_queries, unsigned long long *tau_g_reset,
                                   int list_mod, int queries_per_group, cudaStream_t st, const ReduceOpts &opts,
                                   const Rescore *rs = nullptr, const WaitFlags *wf = nullptr,
                                   unsigned long long *slot_reset = nullptr) {
    ReduceParams<IdT> p;
    p.slot_reset = k_out <= 32 ? slot_reset : nullptr;
    // internal (u32 row id) lists only: the scans emit them sorted; the public merge API does not require it
    p.early_exit = (sizeof(IdT) == 4 && opts.early) ? 1 : 0;
    p.trigger_early = (sizeof(IdT) == 4 && opts.trigger_early) ? 1 : 0;
    p.wait_flags = wf ? wf->flags : nullptr;
    p.wait_n = wf ? wf->n : 0;
    p.wait_epoch = wf ? wf->epoch : 0;
    p.tau_g_reset = tau_g_reset;
    p.rs_rows = rs ? static_cast<const unsigned char *>(rs->rows) : nullptr;
    p.rs_stride = rs ? rs->stride : 0;
    p.rs_dim = rs ? rs->dim : 0;
    p.rs_bf16 = rs ? rs->bf16 : 1;
    p.rs_q = rs ? rs->q : nullptr;
    p.rs_q_stride = rs ? rs->q_stride : 0;
    p.k_final = rs ? rs->k_final : 0;
    p.list_mod = list_mod;
    p.queries_per_group = queries_per_group;
    p.cand_s = cand_s;
    p.cand_i = cand_i;
    p.list_stride = list_stride;
    p.list_stride_i = list_stride_i;
    p.query_stride = query_stride;
    p.n_lists = n_lists;
    p.n_queries = n_queries;
    p.k_in = k_in;
    p.k_out = k_out;
    p.id_base = id_base;
    p.out_s = out_s;
    p.out_i = out_i;
    // programmatic dependent launch: the reduce grid is scheduled while the scan drains and
    // blocks in griddepcontrol.wait until the scan's candidates are visible
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = opts.no_pdl ? 0 : 1;
    if constexpr (sizeof(IdT) == 4) {
        // radix-select reduce (scan.cuh; default on since the round-2 timings, vqa_tuning_t::reduce_select):
        // needs every candidate of a query in shared memory at once and no peer flags to wait for.
        // Taken for k_out > 32, and for the screen-then-rescore reduce of any k_out: one CTA per query re-scores
        // the survivors with coalesced row reads, one warp per candidate (the warp-per-query kernel reads 32 cold
        // rows with one lane each: 63 us per 128 queries on the B200).
        const long long n_cand = (long long)(n_lists / (list_mod > 1 ? list_mod : 1)) * k_in;
        const size_t smem = select_smem_bytes(n_cand);
        if (opts.select && wf == nullptr && smem <= kSelectSmemLimit &&
            (k_out > 32 || (rs != nullptr && p.slot_reset == nullptr))) {
            cudaError_t e = cudaFuncSetAttribute(reduce_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            cfg.gridDim = dim3(n_queries);
            cfg.blockDim = dim3(kSelThreads);
            cfg.dynamicSmemBytes = smem;
            return cudaLaunchKernelEx(&cfg, reduce_select_kernel, p);
        }
    }
    if (k_out <= 32) {
        cfg.gridDim = dim3((n_queries + kReduceWarpsPerCta - 1) / kReduceWarpsPerCta);
        cfg.blockDim = dim3(kReduceWarpsPerCta * 32);
        cfg.dynamicSmemBytes = 0;
        return cudaLaunchKernelEx(&cfg, reduce_topk_warp_kernel<IdT>, p);
    }
    if (rs != nullptr) return cudaErrorInvalidValue;  // only the radix-select reduce re-scores beyond 32 candidates
    cfg.gridDim = dim3(n_queries);
    cfg.blockDim = dim3(kReduceBigWarps * 32);
    cfg.dynamicSmemBytes = list_smem_bytes<IdT>(kReduceBigWarps, k_out);
    return cudaLaunchKernelEx(&cfg, reduce_topk_kernel<IdT>, p);
}

cudaError_t launch_reduce_u32(const float *cand_s, const uint32_t *cand_i, long long list_stride,
                              long long query_stride, int n_lists, int k_in, int k_out, long long id_base,
                              float *out_s, long long *out_i, int n_queries, unsigned long long *tau_g_reset,
                              int list_mod, int queries_per_group, cudaStream_t st, const ReduceOpts &opts,
                              const Rescore *rs, unsigned long long *slot_reset) {
    return launch_reduce_t<uint32_t>(cand_s, cand_i, list_stride, list_stride, query_stride, n_lists, k_in, k_out, id_base, out_s,
                                     out_i, n_queries, tau_g_reset, list_mod, queries_per_group, st, opts, rs, nullptr, slot_reset);
}
cudaError_t launch_reduce_i64(const float *cand_s, const long long *cand_i, long long list_stride,
                              long long list_stride_i, long long query_stride, int n_lists, int k_in, int k_out,
                              long long id_base, float *out_s, long long *out_i, int n_queries, cudaStream_t st,
                              const WaitFlags *wf) {
    return launch_reduce_t<long long>(cand_s, cand_i, list_stride, list_stride_i, query_stride, n_lists, k_in, k_out, id_base,
                                      out_s, out_i, n_queries, nullptr, 1, 1, st, ReduceOpts(), nullptr, wf);
}

cudaError_t launch_exchange_push(const void *local, size_t bytes, void *const *peer_slots,
                                 unsigned long long *const *peer_flags, int world, unsigned long long epoch,
                                 cudaStream_t st) {
    PushParams p;
    p.local = static_cast<const uint4 *>(local);
    p.n16 = bytes / 16;
    for (int r = 0; r < 16; ++r) {
        p.peer_slot[r] = r < world ? static_cast<uint4 *>(peer_slots[r]) : nullptr;
        p.peer_flag[r] = r < world ? peer_flags[r] : nullptr;
    }
    p.epoch = epoch;
    exchange_push_kernel<<<world, 256, 0, st>>>(p);
    return cudaGetLastError();
}

template <typename T, typename MT>
static cudaError_t launch_pool_tm(const void *hidden, const void *mask, int batch, int seq, int dim, int normalize,
                                  float *out, cudaStream_t st) {
    constexpr int E = Elem<T>::E;
    const int nchunks = dim / E;
    const int cw = nchunks < kPoolThreads ? nchunks : kPoolThreads;
    const int G = kPoolThreads / cw;
    const size_t smem = ((size_t)(G + 1) * dim + 32) * sizeof(float);
    auto kern = pool_normalize_kernel<T, MT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<batch, kPoolThreads, smem, st>>>(static_cast<const unsigned char *>(hidden),
                                           static_cast<const MT *>(mask), seq, dim, normalize, out);
    return cudaGetLastError();
}

template <typename T, int ITERS>
static cudaError_t launch_pool_warp(const void *hidden, const void *mask, int m_dtype, int batch, int seq, int dim,
                                    int normalize, float *out, cudaStream_t st) {
    const size_t smem = ((size_t)(kPoolWarps + 1) * dim + 32 + kPoolWarps) * sizeof(float);  // part, pooled, red[32], cntw
    auto kern = pool_normalize_warp_kernel<T, ITERS>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<batch, kPoolFastThreads, smem, st>>>(static_cast<const unsigned char *>(hidden), mask, m_dtype, seq, dim,
                                               normalize, out);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_pool_t(const void *hidden, const void *mask, int m_dtype, int batch, int seq, int dim,
                                 int normalize, float *out, cudaStream_t st) {
    // rows of up to 2 KB (16-bit) / 4 KB (fp32): warp-per-token fast path
    const int iters = (dim / Elem<T>::E + 31) / 32;
    if (iters <= 1) return launch_pool_warp<T, 1>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
    if (iters <= 2) return launch_pool_warp<T, 2>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
    if (iters <= 3) return launch_pool_warp<T, 3>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
    if (iters <= 4) return launch_pool_warp<T, 4>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
    if constexpr (Elem<T>::E == 4) {
        if (iters <= 6) return launch_pool_warp<T, 6>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
        if (iters <= 8) return launch_pool_warp<T, 8>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
    }
    switch (m_dtype) {
        case VQA_I64: return launch_pool_tm<T, long long>(hidden, mask, batch, seq, dim, normalize, out, st);
        case VQA_I32: return launch_pool_tm<T, int>(hidden, mask, batch, seq, dim, normalize, out, st);
        case VQA_U8: return launch_pool_tm<T, unsigned char>(hidden, mask, batch, seq, dim, normalize, out, st);
        case VQA_F32: return launch_pool_tm<T, float>(hidden, mask, batch, seq, dim, normalize, out, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_pool(const void *hidden, int h_dtype, const void *mask, int m_dtype, int batch, int seq, int dim,
                        int normalize, float *out, cudaStream_t st) {
    switch (h_dtype) {
        case VQA_F32: return launch_pool_t<float>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
        case VQA_BF16: return launch_pool_t<__nv_bfloat16>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
        case VQA_F16: return launch_pool_t<__half>(hidden, mask, m_dtype, batch, seq, dim, normalize, out, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_normalize(const float *in, long long in_stride, long long n_rows, int dim, float *out,
                             long long out_stride, void *cast_out, int cast_kind, long long cast_stride,
                             cudaStream_t st) {
    const int rows_per_block = 256 / 32;
    const long long blocks = (n_rows + rows_per_block - 1) / rows_per_block;
    normalize_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, in_stride, n_rows, dim, out, out_stride, cast_out,
                                                           cast_kind, cast_stride);
    return cudaGetLastError();
}

cudaError_t launch_agree(const long long *ids_a, const float *sa, const long long *ids_b, const float *sb,
                         long long n, double threshold, unsigned char *accept, float *combined, cudaStream_t st) {
    const long long blocks = (n + 255) / 256;
    agree_kernel<<<(unsigned)blocks, 256, 0, st>>>(ids_a, sa, ids_b, sb, n, threshold, accept, combined);
    return cudaGetLastError();
}

int segmerge_max_k() { return kWideK; }
int segmerge_max_cand() { return kSegMaxCand; }

cudaError_t launch_merge_segments(const float *seg_s, const long long *seg_i, int n_seg, int n_queries, int k_seg,
                                  int k_out, float *out_s, long long *out_i, int *saturated, cudaStream_t st) {
    SegMergeParams p;
    p.seg_s = seg_s;
    p.seg_i = seg_i;
    p.n_seg = n_seg;
    p.n_queries = n_queries;
    p.k_seg = k_seg;
    p.k_out = k_out;
    p.out_s = out_s;
    p.out_i = out_i;
    p.saturated = saturated;
    int n_pad = 2;
    while (n_pad < n_seg * k_seg) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * sizeof(unsigned long long);
    cudaError_t e = cudaFuncSetAttribute(merge_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(saturated, 0, (size_t)n_seg * sizeof(int), st);
    if (e != cudaSuccess) return e;
    merge_segments_kernel<<<n_queries, kSegThreads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace vqa
