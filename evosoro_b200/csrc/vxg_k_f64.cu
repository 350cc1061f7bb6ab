// fp64 instantiation (logic-parity / high-accuracy mode)
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f64(const BatchDev<double>& B, const LaunchCfg& c, bool smem) {
    return smem ? launch_by_stride<double, true, false, false, false>(B, c) : launch_by_stride<double, false, false, false, false>(B, c);
}
}
