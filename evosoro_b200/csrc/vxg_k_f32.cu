// production kernels: fp32, shared-memory resident, compact bond table; plus the stats kernels
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f32_compact(const BatchDev<float>& B, const LaunchCfg& c) { return launch_by_stride<float, true, true, false, false>(B, c); }
void launch_stats_f32(const BatchDev<float>& B, StatsDev* out, int n_robots, cudaStream_t s) { vx_stats_kernel<float><<<n_robots, 256, 0, s>>>(B, out); }
void launch_stats_f64(const BatchDev<double>& B, StatsDev* out, int n_robots, cudaStream_t s) { vx_stats_kernel<double><<<n_robots, 256, 0, s>>>(B, out); }
}
