// production kernels: fp32, shared-memory resident, compact bond table; plus the stats kernels
#include <cstdlib>
#include <cstring>
#include "vxg_launch.cuh"
#include "vxg_gather.cuh"
namespace vx {
// The gather form (vxg_gather.cuh: damping history in tensor memory, shares gathered by the -axis neighbour) runs every stride
// that has one voxel per thread.  VXG_KERNEL_FORM selects another form of the same arithmetic for A/B measurements and
// bit-for-bit comparisons (read per launch: the tests flip it inside one process):
//   cluster : stride-1024 robots split over a two-CTA thread-block cluster (vxg_cluster.cuh: halves of two different robots share an
//          SM, the halves of one robot meet at the hardware cluster barrier and read each other's boundary voxels through distributed
//          shared memory).  Bit-identical, and 3-7 % SLOWER than the single-CTA kernel (10^3 x 1024: 1.91e10 relaxed arrive / 1.85e10
//          release-acquire vs 1.97e10): two barrier domains of 16 warps do not issue more than one of 32 -- ncu: 70.9 % issue-active
//          against 72.9 %, 4 % more instructions, and 28 % of all warp samples waiting at the cluster barrier for the partner CTA, whose
//          progress depends on what it shares ITS SM with (profiles/r02_ncu_cluster10.txt).  What two 8^3 robots per SM gained came
//          from doubling the resident warps, not from splitting the barrier domain.  Kept selectable as the record of that experiment.
//   push : the push form (vxg_kernels.cuh: shares pushed into shared-memory accumulators, four barriers per step)
//   flow : the gather form as a per-warp dataflow (phase mbarriers between neighbouring warps instead of CTA-wide barriers).
//          Measured 6-12 % SLOWER than the barrier form (10^3: 1.72e10 vs 1.83e10, 8^3: 1.79e10 vs 2.03e10): a warp and the
//          warps its voxels' neighbours live in depend on each other in BOTH directions every phase, so the whole CTA still
//          advances in lock step and the mbarrier waits only add instructions.  Kept selectable as the record of that experiment.
cudaError_t launch_f32_compact(const BatchDev<float>& B, const LaunchCfg& c) {
    const char* form = getenv("VXG_KERNEL_FORM");
    const bool push = form && !strcmp(form, "push"), flow = form && !strcmp(form, "flow");
    if (c.nv == 1024 && form && !strcmp(form, "cluster")) { cudaError_t err = cudaSuccess; if (launch_f32_cluster(B, c, &err) == 0) return err; }
    if (!push && vxg_use_rt(c.nv)) return launch_gather<0>(B, c, flow);
    if (!push) switch (c.nv) {
        case 64: return launch_gather<64>(B, c, flow);
        case 96: return launch_gather<96>(B, c, flow);
        case 128: return launch_gather<128>(B, c, flow);
        case 160: return launch_gather<160>(B, c, flow);
        case 192: return launch_gather<192>(B, c, flow);
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
