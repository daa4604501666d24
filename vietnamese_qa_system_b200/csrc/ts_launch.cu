// ts_launch.cu -- instantiation + launch of the TMEM-resident-query tcgen05 kernel (ts.cuh).
#include "launch.h"
#include "ts.cuh"

namespace vqa {

template <bool BF16, int KL, bool QS, bool M64 = false>
static cudaError_t launch_ts_tk(const TsLaunch &a, cudaStream_t st) {
    TsParams p;
    p.q = a.q;
    p.q_stride = a.q_stride;
    p.nq = a.nq;
    p.per_cta = (a.split || M64) ? 64 : 128;
    p.split = a.split;
    p.a_fp16 = a.a_fp16;
    p.k = a.k;
    p.n_rows = a.n_rows;
    p.dim = a.dim;
    p.cand_s = a.cand_s;
    p.cand_i = a.cand_i;
    p.cand_stride = a.cand_stride;
    p.n_tiles = (int)((a.n_rows + kTsDocs - 1) / kTsDocs);
    p.n_stages = a.stages;
    p.kps = a.kps;
    p.tma_policy = a.n_groups > 1 && !a.multicast ? 0x1000000000000000ull : 0x12F0000000000000ull;
    p.tau_g = a.tau_g;
    p.epoch = a.epoch;
    p.n_groups = a.n_groups;
    p.multicast = a.multicast;
    p.ks = QS ? a.ks : 0;
    p.timeline = a.timeline;
    const size_t smem = ts_smem_bytes_rt(a.k, a.stages * a.kps, a.split, p.ks, a.nq, QS ? 1 : 0, M64 ? 1 : 0);
    auto kern = ts_topk_kernel<BF16, KL, QS, M64>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (!a.multicast && !a.pdl) {
        kern<<<a.grid, kMmaThreads, smem, st>>>(*a.tmap, p);
        return cudaGetLastError();
    }
    // cluster launch (TMA multicast) and / or programmatic dependent launch: a.pdl marks the 2nd, 3rd ... scan of
    // ONE search, which depends on nothing the previous launch's reduce kernel produces, so it may start while
    // that reduce is still running (the kernel never executes griddepcontrol.wait)
    cudaLaunchAttribute attr[2];
    unsigned n_attr = 0;
    if (a.multicast) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = (unsigned)a.n_groups;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    if (a.pdl) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.grid);
    cfg.blockDim = dim3(kMmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    return cudaLaunchKernelEx(&cfg, kern, *a.tmap, p);
}

template <bool BF16, bool QS, bool M64 = false>
static cudaError_t launch_ts_t(const TsLaunch &a, cudaStream_t st) {
    switch (ts_reg_list_len(a.k)) {
        case 16: return launch_ts_tk<BF16, 16, QS, M64>(a, st);
        case 32: return launch_ts_tk<BF16, 32, QS, M64>(a, st);
        default: return launch_ts_tk<BF16, 0, QS, M64>(a, st);
    }
}

cudaError_t launch_ts(const TsLaunch &a, cudaStream_t st) {
    const int kb = a.dim / kBlockK;
    if (a.qs) {
        // the TMEM part of the query block must leave at least one accumulator stage
        if (a.ks < 0 || a.ks > kb || (kb - a.ks) * (kBlockK / 2) + kTsDocs > 512) return cudaErrorInvalidValue;
        if (a.m64) {   // M = 64 instructions: one chunk of <= 64 queries, no hi/lo rows, no cluster
            if (a.split || a.nq > 64 || a.n_groups != 1) return cudaErrorInvalidValue;
            return a.bf16 ? launch_ts_t<true, true, true>(a, st) : launch_ts_t<false, true, true>(a, st);
        }
        return a.bf16 ? launch_ts_t<true, true>(a, st) : launch_ts_t<false, true>(a, st);
    }
    if (a.dim / 2 + kTsDocs > 512) return cudaErrorInvalidValue;
    return a.bf16 ? launch_ts_t<true, false>(a, st) : launch_ts_t<false, false>(a, st);
}

size_t ts_smem_bytes(int k, int boxes, int split, int ks, int nq, int qs, int m64) {
    return ts_smem_bytes_rt(k, boxes, split, ks, nq, qs, m64);
}

}  // namespace vqa
