// scan_launch.cuh -- template dispatch for the streaming scan kernel (one
// translation unit per storage type includes this).
#pragma once
#include "launch.h"
#include "scan.cuh"

namespace vqa {

template <typename T, int BT, int R, int ITERS>
cudaError_t launch_scan_one(const ScanLaunch &a, cudaStream_t st) {
    ScanParams p;
    p.rows = static_cast<const unsigned char *>(a.rows);
    p.n_rows = a.n_rows;
    p.row_stride_bytes = a.row_stride_bytes;
    p.dim = a.dim;
    p.q = a.q;
    p.q_stride = a.q_stride;
    p.nq = a.nq;
    p.k = a.k;
    p.cand_s = a.cand_s;
    p.cand_i = a.cand_i;
    p.cand_stride = a.cand_stride;
    const size_t smem = scan_smem_bytes<T, BT>(a.dim, a.k);
    auto kern = scan_topk_kernel<T, BT, R, ITERS>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int passes = (a.nq + BT - 1) / BT;
    kern<<<dim3((unsigned)a.grid, (unsigned)passes), kScanThreads, smem, st>>>(p);
    return cudaGetLastError();
}

// I1, I2: the two fully-unrolled row lengths (in 512-byte warp iterations) this
// storage type is specialised for; anything else takes the generic loop.
template <typename T, int BT, int R, int I1, int I2>
cudaError_t launch_scan_bt(const ScanLaunch &a, cudaStream_t st) {
    const long long row_bytes = (long long)a.dim * (long long)sizeof(T);
    if (row_bytes == (long long)I1 * 512) return launch_scan_one<T, BT, R, I1>(a, st);
    if (row_bytes == (long long)I2 * 512) return launch_scan_one<T, BT, R, I2>(a, st);
    return launch_scan_one<T, BT, R, 0>(a, st);
}

template <typename T, int R, int I1, int I2>
cudaError_t launch_scan_t(const ScanLaunch &a, cudaStream_t st) {
    switch (a.bt) {
        case 1: return launch_scan_bt<T, 1, R, I1, I2>(a, st);
        case 2: return launch_scan_bt<T, 2, R, I1, I2>(a, st);
        case 4: return launch_scan_bt<T, 4, R, I1, I2>(a, st);
        case 8: return launch_scan_bt<T, 8, R, I1, I2>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace vqa
