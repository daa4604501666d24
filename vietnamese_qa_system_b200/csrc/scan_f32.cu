// fp32 rows (verify mode's native storage): 2 rows per warp step; unrolled for
// dim 384 (3 x 512 B) and dim 768 (6 x 512 B).
#include "scan_launch.cuh"
namespace vqa {
cudaError_t launch_scan_f32(const ScanLaunch &a, cudaStream_t st) { return launch_scan_t<float, 2, 3, 6>(a, st); }
}  // namespace vqa
