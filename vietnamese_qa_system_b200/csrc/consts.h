// consts.h -- compile-time constants shared by host dispatch and device code.
#pragma once
#include <stddef.h>

namespace vqa {

constexpr int kMaxK = 128;                            // fused top-k lists: up to 4 entries per lane
constexpr int kMaxKpl = kMaxK / 32;

// tensor-core path tile geometry
constexpr int kMmaThreads = 192;
constexpr int kTileRows = 128;                        // documents per MMA tile (UMMA M)
constexpr int kBlockK = 64;                           // elements per k-block (128 bytes, one swizzle row)
constexpr int kStageBytes = kTileRows * kBlockK * 2;  // 16 KB per TMA stage
constexpr int kMaxStages = 12;
constexpr int kMaxAccStages = 8;

inline size_t list_bytes_rt(int nq, int k, int id_bytes) {
    int kcap = ((k + 31) / 32) * 32;
    return (size_t)nq * kcap * (4 + id_bytes) + (size_t)nq * 8;
}
// shared memory of the tensor-core kernel: [align slack][Q tiles (hi|lo columns)][A ring][barriers][lists]
// `ncol` = MMA N (2 x queries per CTA with hi/lo columns, 1 x in screen mode); `boxes` = number of
// 16 KB document boxes in the ring
inline size_t mma_smem_bytes_rt(int ncol, int dim, int k, int boxes, int split = 1) {
    const int nq = split ? ncol / 2 : ncol;
    // k <= 32 and <= 32 queries per CTA: the lists live in registers, no shared-memory lists at all
    const size_t lists = (nq <= 32 && k <= 32) ? 0 : list_bytes_rt(nq, k, 4);
    return 1024 + (size_t)(dim / kBlockK) * ncol * 128 + (size_t)boxes * kStageBytes + 1024 + lists;
}

}  // namespace vqa
