// bf16 rows: 4 rows per warp step; unrolled for dim 768 (3 x 512 B) and 1024 (4 x 512 B).
#include "scan_launch.cuh"
namespace vqa {
cudaError_t launch_scan_bf16(const ScanLaunch &a, cudaStream_t st) {
    return launch_scan_t<__nv_bfloat16, 4, 3, 4>(a, st);
}
}  // namespace vqa
