// pair_launch.cu -- instantiation + launch of the CTA-pair (cta_group::2) tcgen05 kernel (pair.cuh).
#include "launch.h"
#include "pair.cuh"

namespace vqa {

template <bool BF16, int KL>
static cudaError_t launch_pair_tk(const PairLaunch &a, cudaStream_t st) {
    PairParams p;
    p.q = a.q;
    p.q_stride = a.q_stride;
    p.nq = a.nq;
    p.k = a.k;
    p.n_rows = a.n_rows;
    p.dim = a.dim;
    p.cand_s = a.cand_s;
    p.cand_i = a.cand_i;
    p.cand_stride = a.cand_stride;
    p.n_tiles = (int)((a.n_rows + kPairDocs - 1) / kPairDocs);
    p.n_stages = a.stages;
    p.kps = a.kps;
    p.tma_policy = 0x12F0000000000000ull;  // evict-first: every document byte is read once per launch
    p.tau_g = a.tau_g;
    p.epoch = a.epoch;
    p.ks = a.ks;
    p.timeline = a.timeline;
    const size_t smem = pair_smem_bytes_rt(a.stages * a.kps, a.ks);
    auto kern = ts_pair_topk_kernel<BF16, KL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<a.grid, kMmaThreads, smem, st>>>(*a.tmap, p);  // __cluster_dims__(2, 1, 1): grid is even
    return cudaGetLastError();
}

cudaError_t launch_pair(const PairLaunch &a, cudaStream_t st) {
    const int kb = a.dim / kBlockK;
    // the tensor-memory part of the query block must leave at least two 128-column accumulator stages
    if (a.grid < 2 || a.grid % 2 != 0 || a.ks < 0 || a.ks > kb || (kb - a.ks) * (kBlockK / 2) + 2 * kPairDocs > 512 ||
        a.k > 32 || a.nq > 2 * kPairRows)
        return cudaErrorInvalidValue;
    if (a.k <= 16) return a.bf16 ? launch_pair_tk<true, 16>(a, st) : launch_pair_tk<false, 16>(a, st);
    return a.bf16 ? launch_pair_tk<true, 32>(a, st) : launch_pair_tk<false, 32>(a, st);
}

size_t pair_smem_bytes(int boxes, int ks) { return pair_smem_bytes_rt(boxes, ks); }

}  // namespace vqa
