// fp32 with self-collision on the production form (what evosoro's GA runs by default: SelfColEnabled = 1 on palette robots): the
// gather form (vxg_gather.cuh) for every stride with one voxel per thread, else the push form with the compact bond table
#include <cstdlib>
#include <cstring>
#include "vxg_launch.cuh"
#include "vxg_gather.cuh"
namespace vx {
cudaError_t launch_f32_collide_compact(const BatchDev<float>& B, const LaunchCfg& c) {
    const char* form = getenv("VXG_KERNEL_FORM");
    if (!(form && !strcmp(form, "push"))) {
        cudaError_t e = cudaSuccess; int miss = 1;
        if (vxg_use_rt(c.nv)) miss = launch_gather_collide<0>(B, c, &e);
        else switch (c.nv) {
            case 64: miss = launch_gather_collide<64>(B, c, &e); break;
            case 96: miss = launch_gather_collide<96>(B, c, &e); break;
            case 128: miss = launch_gather_collide<128>(B, c, &e); break;
            case 160: miss = launch_gather_collide<160>(B, c, &e); break;
            case 192: miss = launch_gather_collide<192>(B, c, &e); break;
            case 224: miss = launch_gather_collide<224>(B, c, &e); break;
            case 256: miss = launch_gather_collide<256>(B, c, &e); break;
            case 352: miss = launch_gather_collide<352>(B, c, &e); break;
            case 512: miss = launch_gather_collide<512>(B, c, &e); break;
            case 736: miss = launch_gather_collide<736>(B, c, &e); break;
            case 1024: miss = launch_gather_collide<1024>(B, c, &e); break;
        }
        if (!miss) return e;
    }
    return launch_by_stride<float, true, true, false, true>(B, c);
}
}
