// land/water layouts: fluid drag on the surface mesh, with or without self-collision
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f32_fluid(const BatchDev<float>& B, const LaunchCfg& c, bool collide) {
    return collide ? launch_by_stride<float, true, false, true, true>(B, c) : launch_by_stride<float, true, false, true, false>(B, c);
}
}
