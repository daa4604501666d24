// mma_launch.cu -- instantiations + launch of the tcgen05 tensor-core kernel.
#include "launch.h"
#include "mma.cuh"


namespace vqa {

// L2 policy for the document stream (vqa_tuning_t::tma_hint: 0 normal, 1 evict-first, 2 evict-last)
static unsigned long long tma_policy(int v) {
    return v == 0 ? 0x1000000000000000ull : (v == 2 ? 0x14F0000000000000ull : 0x12F0000000000000ull);
}

template <bool BF16, int NCOL, bool SPLIT>
static cudaError_t launch_mma_one(const MmaLaunch &a, cudaStream_t st) {
    MmaParams p;
    p.q = a.q;
    p.q_stride = a.q_stride;
    p.nq = a.nq;
    p.n_groups = a.n_groups;
    p.k = a.k;
    p.n_rows = a.n_rows;
    p.dim = a.dim;
    // fp16's 11-bit significand leaves the residual near its subnormal range: scale it up (exactly)
    p.lo_scale = BF16 ? 1.0f : 2048.0f;
    p.lo_inv_scale = BF16 ? 1.0f : 1.0f / 2048.0f;
    p.cand_s = a.cand_s;
    p.cand_i = a.cand_i;
    p.cand_stride = a.cand_stride;
    p.n_tiles = (int)((a.n_rows + kTileRows - 1) / kTileRows);
    p.n_stages = a.stages;
    p.kps = a.kps;
    p.tau_g = a.tau_g;
    p.epoch = a.epoch;
    // side-by-side chunks re-read each tile from L2: keep it there (normal policy) instead of evict-first
    p.tma_policy = a.n_groups > 1 ? 0x1000000000000000ull : tma_policy(a.tma_hint);
    p.multicast = a.multicast;
    p.timeline = a.timeline;
    p.slot_g = ((SPLIT ? NCOL / 2 : NCOL) <= 32 && a.k <= 32) ? a.slot_g : nullptr;  // register-list path only
    p.tile_ctr = (a.n_groups == 1 && !a.multicast) ? a.tile_ctr : nullptr;
    const size_t smem = mma_smem_bytes_rt(NCOL, a.dim, a.k, a.stages * a.kps, SPLIT ? 1 : 0);
    auto kern = mma_topk_kernel<BF16, NCOL, SPLIT>;
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

template <bool BF16, int NCOL, bool SPLIT>
static int max_clusters_one(int cluster, size_t smem) {
    auto kern = mma_topk_kernel<BF16, NCOL, SPLIT>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cluster * 64));
    cfg.blockDim = dim3(kMmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

template <bool BF16>
static int max_clusters_t(int ncol, int split, int cluster, size_t smem) {
    if (!split) {
        switch (ncol) {
            case 16: return max_clusters_one<BF16, 16, false>(cluster, smem);
            case 32: return max_clusters_one<BF16, 32, false>(cluster, smem);
            default: return 0;
        }
    }
    switch (ncol) {
        case 16: return max_clusters_one<BF16, 16, true>(cluster, smem);
        case 32: return max_clusters_one<BF16, 32, true>(cluster, smem);
        case 64: return max_clusters_one<BF16, 64, true>(cluster, smem);
        case 128: return max_clusters_one<BF16, 128, true>(cluster, smem);
        default: return 0;
    }
}

int mma_max_active_clusters(bool bf16, int ncol, int split, int cluster, size_t smem_bytes) {
    return bf16 ? max_clusters_t<true>(ncol, split, cluster, smem_bytes)
                : max_clusters_t<false>(ncol, split, cluster, smem_bytes);
}

template <bool BF16>
static cudaError_t launch_mma_t(const MmaLaunch &a, cudaStream_t st) {
    if (!a.split) {
        switch (a.ncol) {
            case 16: return launch_mma_one<BF16, 16, false>(a, st);
            case 32: return launch_mma_one<BF16, 32, false>(a, st);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (a.ncol) {
        case 16: return launch_mma_one<BF16, 16, true>(a, st);
        case 32: return launch_mma_one<BF16, 32, true>(a, st);
        case 64: return launch_mma_one<BF16, 64, true>(a, st);
        case 128: return launch_mma_one<BF16, 128, true>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_mma(const MmaLaunch &a, cudaStream_t st) {
    return a.bf16 ? launch_mma_t<true>(a, st) : launch_mma_t<false>(a, st);
}

}  // namespace vqa
