// fp32 with self-collision (contact-bond list rebuild, contact forces, max-velocity reduction)
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f32_collide(const BatchDev<float>& B, const LaunchCfg& c) { return launch_by_stride<float, true, false, false, true>(B, c); }
}
