// fp32 with self-collision on the production form: compact bond table, feature-stripped like vxg_k_f32.cu (what evosoro's GA
// runs by default: SelfColEnabled = 1 on palette robots)
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f32_collide_compact(const BatchDev<float>& B, const LaunchCfg& c) { return launch_by_stride<float, true, true, false, true>(B, c); }
}
