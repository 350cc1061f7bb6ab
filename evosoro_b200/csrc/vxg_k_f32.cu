// production kernels: fp32, shared-memory resident, compact bond table; plus the stats kernels
#include <cstdlib>
#include <cstring>
#include "vxg_launch.cuh"
#include "vxg_gather.cuh"
namespace vx {
// class-table capacities the step loop is specialised for (palette robots: <= 8 voxel classes, <= 63 bond classes after
// re-keying by voxel-class pair); larger tables run the instantiation with run-time capacities
#define VXG_CV 8
#define VXG_CB 64
template <int NV, bool FLOW, int CV, int CB> static cudaError_t launch_gather_form(const BatchDev<float>& B, const LaunchCfg& c) {
    auto kern = vx_step_kernel_g<NV, FLOW, CV, CB>;
    const size_t smem = gather_smem_bytes(NV, CV > 0 ? CV : B.max_vclass, CB > 0 ? CB : B.max_bclass + 1, FLOW);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<c.n_robots, NV, smem, c.stream>>>(B, c.n_steps);
    return cudaGetLastError();
}
// does a CTA of this form fit next to the others the launch bounds ask for (227 KB per CTA, 228 KB per SM with 1 KB reserved per CTA)?
template <int NV> static bool gather_fits(int mv, int mb, bool flow) {
    const size_t per_cta = gather_smem_bytes(NV, mv, mb, flow);
    const int ctas = vx_min_ctas(NV, NV, false, false);
    return per_cta <= 227u * 1024u && (per_cta + 1024u) * (size_t)ctas <= 228u * 1024u;
}
template <int NV> static cudaError_t launch_gather(const BatchDev<float>& B, const LaunchCfg& c, bool flow) {
    const bool fixed = B.max_vclass <= VXG_CV && B.max_bclass + 1 <= VXG_CB && gather_fits<NV>(VXG_CV, VXG_CB, false);
    if (flow) {   // experiment form, run-time capacities only
        if (gather_fits<NV>(B.max_vclass, B.max_bclass + 1, true)) return launch_gather_form<NV, true, 0, 0>(B, c);
    }
    return fixed ? launch_gather_form<NV, false, VXG_CV, VXG_CB>(B, c) : launch_gather_form<NV, false, 0, 0>(B, c);
}
// The gather form (vxg_gather.cuh: damping history in tensor memory, shares gathered by the -axis neighbour) runs every stride
// that has one voxel per thread.  VXG_KERNEL_FORM selects another form of the same arithmetic for A/B measurements and
// bit-for-bit comparisons (read per launch: the tests flip it inside one process):
//   push : the push form (vxg_kernels.cuh: shares pushed into shared-memory accumulators, four barriers per step)
//   flow : the gather form as a per-warp dataflow (phase mbarriers between neighbouring warps instead of CTA-wide barriers).
//          Measured 6-12 % SLOWER than the barrier form (10^3: 1.72e10 vs 1.83e10, 8^3: 1.79e10 vs 2.03e10): a warp and the
//          warps its voxels' neighbours live in depend on each other in BOTH directions every phase, so the whole CTA still
//          advances in lock step and the mbarrier waits only add instructions.  Kept selectable as the record of that experiment.
cudaError_t launch_f32_compact(const BatchDev<float>& B, const LaunchCfg& c) {
    const char* form = getenv("VXG_KERNEL_FORM");
    const bool push = form && !strcmp(form, "push"), flow = form && !strcmp(form, "flow");
    if (!push) switch (c.nv) {
        case 64: return launch_gather<64>(B, c, flow);
        case 128: return launch_gather<128>(B, c, flow);
        case 224: return launch_gather<224>(B, c, flow);
        case 256: return launch_gather<256>(B, c, flow);
        case 352: return launch_gather<352>(B, c, flow);
        case 512: return launch_gather<512>(B, c, flow);
        case 736: return launch_gather<736>(B, c, flow);
        case 1024: return launch_gather<1024>(B, c, flow);
    }
    return launch_by_stride<float, true, true, false, false>(B, c);
}
void launch_stats_f32(const BatchDev<float>& B, StatsDev* out, int n_robots, cudaStream_t s) { vx_stats_kernel<float><<<n_robots, 256, 0, s>>>(B, out); }
void launch_stats_f64(const BatchDev<double>& B, StatsDev* out, int n_robots, cudaStream_t s) { vx_stats_kernel<double><<<n_robots, 256, 0, s>>>(B, out); }
}
