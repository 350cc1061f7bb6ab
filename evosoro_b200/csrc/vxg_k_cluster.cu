// production kernel for stride-1024 robots: one robot per two-CTA cluster (vxg_cluster.cuh)
#include "vxg_launch.cuh"
#include "vxg_cluster.cuh"
namespace vx {
int launch_f32_cluster(const BatchDev<float>& B, const LaunchCfg& c, cudaError_t* err) { return launch_cluster(B, c, err); }
}
