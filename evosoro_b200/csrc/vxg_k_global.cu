// fp32 self-collision / land-water layouts on the blob in GLOBAL memory (wide tables in shared memory): robots whose blob, mesh
// arrays, contact list and class tables do not fit in shared memory together -- e.g. an evolved swimmer with a class per voxel
// (per-voxel phase offset AND stiffness: sample_vxa/full_cross.vxa has 339 voxel and 706 bond classes for 343 voxels)
#include "vxg_launch.cuh"
namespace vx {
cudaError_t launch_f32_extras_global(const BatchDev<float>& B, const LaunchCfg& c, bool fluid, bool collide) {
    if (fluid) return collide ? launch_by_stride<float, false, false, true, true>(B, c) : launch_by_stride<float, false, false, true, false>(B, c);
    return launch_by_stride<float, false, false, false, true>(B, c);
}
}
