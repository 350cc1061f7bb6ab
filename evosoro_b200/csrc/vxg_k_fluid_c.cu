// land/water layouts on the production form (compact bond table, feature-stripped), with or without self-collision
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f32_fluid_compact(const BatchDev<float>& B, const LaunchCfg& c, bool collide) {
    return collide ? launch_by_stride<float, true, true, true, true>(B, c) : launch_by_stride<float, true, true, true, false>(B, c);
}
}
