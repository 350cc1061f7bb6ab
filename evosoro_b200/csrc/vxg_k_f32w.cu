// fp32 with the wide bond table: shared-memory resident or on the blob in global memory
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f32_wide(const BatchDev<float>& B, const LaunchCfg& c, bool smem) {
    return smem ? launch_by_stride<float, true, false, false, false>(B, c) : launch_by_stride<float, false, false, false, false>(B, c);
}
}
