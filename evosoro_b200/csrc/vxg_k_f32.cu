// production kernels: fp32, shared-memory resident, compact bond table; plus the stats kernels
#include <cstdlib>
#include "vxg_launch.cuh"
#include "vxg_gather.cuh"
namespace vx {
template <int NV> static cudaError_t launch_gather(const BatchDev<float>& B, const LaunchCfg& c) {
    auto kern = vx_step_kernel_g<NV>;
    const size_t smem = gather_smem_bytes(NV, B.max_vclass, B.max_bclass);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<c.n_robots, NV, smem, c.stream>>>(B, c.n_steps);
    return cudaGetLastError();
}
// The gather form (vxg_gather.cuh: damping history in tensor memory, two barriers per timestep) runs every stride that has one
// voxel per thread; VXG_PUSH_KERNEL=1 selects the push form (vxg_kernels.cuh) for A/B measurements and bit-for-bit comparisons.
cudaError_t launch_f32_compact(const BatchDev<float>& B, const LaunchCfg& c) {
    const bool push = getenv("VXG_PUSH_KERNEL") != nullptr;   // read per launch: the tests flip it inside one process
    if (!push) switch (c.nv) {
        case 64: return launch_gather<64>(B, c);
        case 128: return launch_gather<128>(B, c);
        case 224: return launch_gather<224>(B, c);
        case 256: return launch_gather<256>(B, c);
        case 352: return launch_gather<352>(B, c);
        case 512: return launch_gather<512>(B, c);
        case 736: return launch_gather<736>(B, c);
        case 1024: return launch_gather<1024>(B, c);
    }
    return launch_by_stride<float, true, true, false, false>(B, c);
}
void launch_stats_f32(const BatchDev<float>& B, StatsDev* out, int n_robots, cudaStream_t s) { vx_stats_kernel<float><<<n_robots, 256, 0, s>>>(B, out); }
void launch_stats_f64(const BatchDev<double>& B, StatsDev* out, int n_robots, cudaStream_t s) { vx_stats_kernel<double><<<n_robots, 256, 0, s>>>(B, out); }
}
