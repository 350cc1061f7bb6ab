// fp64 self-collision / land-water layouts on the blob in global memory (logic-parity mode)
#include "vxg_launch.cuh"
namespace vx {
// the logic-parity (fp64) instantiation of the same: the reference's arithmetic type for the fluid-drag and self-collision phases
cudaError_t launch_f64_extras_global(const BatchDev<double>& B, const LaunchCfg& c, bool fluid, bool collide) {
    if (fluid) return collide ? launch_by_stride<double, false, false, true, true>(B, c) : launch_by_stride<double, false, false, true, false>(B, c);
    return launch_by_stride<double, false, false, false, true>(B, c);
}
}
