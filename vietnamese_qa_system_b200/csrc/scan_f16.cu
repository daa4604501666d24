// fp16 rows: 4 rows per warp step; unrolled for dim 768 and 1024.
#include "scan_launch.cuh"
namespace vqa {
cudaError_t launch_scan_f16(const ScanLaunch &a, cudaStream_t st) { return launch_scan_t<__half, 4, 3, 4>(a, st); }
}  // namespace vqa
