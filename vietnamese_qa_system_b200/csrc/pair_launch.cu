// pair_launch.cu -- instantiation + launch of the CTA-pair (cta_group::2) tcgen05 kernel (pair.cuh).
#include "launch.h"
#include "pair.cuh"

namespace vqa {

template <bool BF16, int KL, bool PAIR>
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
    const size_t smem = pair_smem_bytes_rt(a.stages * a.kps * (PAIR ? 1 : 2), a.ks);
    auto kern = ts_pair_topk_kernel<BF16, KL, PAIR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (!PAIR) {
        kern<<<a.grid, kMmaThreads, smem, st>>>(*a.tmap, p);
        return cudaGetLastError();
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;   // the CTA pair: two SMs of one TPC
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.grid);
    cfg.blockDim = dim3(kMmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, *a.tmap, p);
}

template <bool BF16, bool PAIR>
static cudaError_t launch_pair_t(const PairLaunch &a, cudaStream_t st) {
    return a.k <= 16 ? launch_pair_tk<BF16, 16, PAIR>(a, st) : launch_pair_tk<BF16, 32, PAIR>(a, st);
}

cudaError_t launch_pair(const PairLaunch &a, cudaStream_t st) {
    const int kb = a.dim / kBlockK;
    // the tensor-memory part of the query block must leave at least two 128-column accumulator stages
    if (a.grid < 1 || a.ks < 0 || a.ks > kb || (kb - a.ks) * (kBlockK / 2) + 2 * kPairDocs > 512 || a.k > 32)
        return cudaErrorInvalidValue;
    if (a.pair) {
        if (a.grid % 2 != 0 || a.nq > 2 * kPairRows) return cudaErrorInvalidValue;
        return a.bf16 ? launch_pair_t<true, true>(a, st) : launch_pair_t<false, true>(a, st);
    }
    if (a.nq > kPairRows) return cudaErrorInvalidValue;
    return a.bf16 ? launch_pair_t<true, false>(a, st) : launch_pair_t<false, false>(a, st);
}

size_t pair_smem_bytes(int boxes, int ks) { return pair_smem_bytes_rt(boxes, ks); }

}  // namespace vqa
