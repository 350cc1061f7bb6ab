// vxg_launch.cuh -- launch plumbing shared by the kernel translation units.  Each vxg_k_*.cu instantiates one
// family of vx_step_kernel (so the families compile in parallel) and exports one plain function that picks
// the instantiation for the batch's voxel stride.
#pragma once
#include <cuda_runtime.h>
#include "vxg_kernels.cuh"

namespace vx {

struct LaunchCfg {
    int n_robots, nv;
    size_t smem_bytes;
    cudaStream_t stream;
    int n_steps, record;
    int vpt;   // voxels per thread the batch runs with (stride 1024 has two instantiations in the production family)
};

template <typename Real, int NV, int VPT, bool SMEM, bool COMPACT, bool FLUID, bool COLLIDE>
inline cudaError_t launch_one(const BatchDev<Real>& B, const LaunchCfg& c) {
    auto kern = vx_step_kernel<Real, NV, VPT, SMEM, COMPACT, FLUID, COLLIDE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<c.n_robots, NV / VPT, c.smem_bytes, c.stream>>>(B, c.n_steps, c.record);
    return cudaGetLastError();
}

// strides the kernels are instantiated for (stride_for); <= 512 threads per CTA
template <typename Real, bool SMEM, bool COMPACT, bool FLUID, bool COLLIDE>
inline cudaError_t launch_by_stride(const BatchDev<Real>& B, const LaunchCfg& c) {
    switch (c.nv) {
        case 64: return launch_one<Real, 64, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 96: return launch_one<Real, 96, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 128: return launch_one<Real, 128, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 160: return launch_one<Real, 160, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 192: return launch_one<Real, 192, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 224: return launch_one<Real, 224, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 256: return launch_one<Real, 256, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 352: return launch_one<Real, 352, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 512: return launch_one<Real, 512, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 736: return launch_one<Real, 736, 1, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 1024:
            if constexpr (COMPACT && SMEM && !FLUID && !COLLIDE && sizeof(Real) == 4) { if (c.vpt == 1) return launch_one<Real, 1024, 1, true, true, false, false>(B, c); }
            return launch_one<Real, 1024, 2, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        case 2048: return launch_one<Real, 2048, 4, SMEM, COMPACT, FLUID, COLLIDE>(B, c);
        // above 2048 voxels: blob in global memory, wide tables, 512 threads x 16 / 64 voxels (rolled loops)
        case 8192: if constexpr (!SMEM && !COMPACT) return launch_one<Real, 8192, 16, false, false, FLUID, COLLIDE>(B, c); break;
        case 32768: if constexpr (!SMEM && !COMPACT) return launch_one<Real, 32768, 64, false, false, FLUID, COLLIDE>(B, c); break;
    }
    return cudaErrorInvalidValue;
}

// one per translation unit
cudaError_t launch_f32_compact(const BatchDev<float>& B, const LaunchCfg& c);                 // production: fp32, smem, compact table
int launch_f32_cluster(const BatchDev<float>& B, const LaunchCfg& c, cudaError_t* err);       // production, stride 1024: one robot per two-CTA cluster (0 = launched)
cudaError_t launch_f32_wide(const BatchDev<float>& B, const LaunchCfg& c, bool smem);         // fp32, wide table, smem or global blob
cudaError_t launch_f64(const BatchDev<double>& B, const LaunchCfg& c, bool smem);             // logic-parity mode
cudaError_t launch_f32_collide(const BatchDev<float>& B, const LaunchCfg& c);                 // + self-collision
cudaError_t launch_f32_collide_compact(const BatchDev<float>& B, const LaunchCfg& c);         // + self-collision, production form
cudaError_t launch_f32_fluid(const BatchDev<float>& B, const LaunchCfg& c, bool collide);     // land/water layouts (+ self-collision)
cudaError_t launch_f32_fluid_compact(const BatchDev<float>& B, const LaunchCfg& c, bool collide);   // the same on the production form
cudaError_t launch_f32_extras_global(const BatchDev<float>& B, const LaunchCfg& c, bool fluid, bool collide);   // the same on the blob in global memory
cudaError_t launch_f64_extras_global(const BatchDev<double>& B, const LaunchCfg& c, bool fluid, bool collide);  // fp64, blob in global memory
void launch_stats_f32(const BatchDev<float>& B, StatsDev* out, int n_robots, cudaStream_t s);
void launch_stats_f64(const BatchDev<double>& B, StatsDev* out, int n_robots, cudaStream_t s);

}  // namespace vx
