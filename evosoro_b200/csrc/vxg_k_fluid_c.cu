// land/water layouts on the production form (compact bond table, feature-stripped), with or without self-collision: the gather
// form (vxg_gather.cuh) where one voxel per thread and the robot with its surface mesh fit in shared memory, else the push form
#include <cstdlib>
#include <cstring>
#include "vxg_launch.cuh"
#include "vxg_gather.cuh"
namespace vx {
cudaError_t launch_f32_fluid_compact(const BatchDev<float>& B, const LaunchCfg& c, bool collide) {
    const char* form = getenv("VXG_KERNEL_FORM");
    if (!(form && !strcmp(form, "push"))) {
        cudaError_t e = cudaSuccess; int miss = 1;
        if (vxg_use_rt(c.nv)) miss = launch_gather_fluid<0>(B, c, collide, &e);
        else switch (c.nv) {
            case 64: miss = launch_gather_fluid<64>(B, c, collide, &e); break;
            case 96: miss = launch_gather_fluid<96>(B, c, collide, &e); break;
            case 128: miss = launch_gather_fluid<128>(B, c, collide, &e); break;
            case 160: miss = launch_gather_fluid<160>(B, c, collide, &e); break;
            case 192: miss = launch_gather_fluid<192>(B, c, collide, &e); break;
            case 224: miss = launch_gather_fluid<224>(B, c, collide, &e); break;
            case 256: miss = launch_gather_fluid<256>(B, c, collide, &e); break;
            case 352: miss = launch_gather_fluid<352>(B, c, collide, &e); break;
            case 512: miss = launch_gather_fluid<512>(B, c, collide, &e); break;
            case 736: miss = launch_gather_fluid<736>(B, c, collide, &e); break;
            case 1024: miss = launch_gather_fluid<1024>(B, c, collide, &e); break;
        }
        if (!miss) return e;
    }
    return collide ? launch_by_stride<float, true, true, true, true>(B, c) : launch_by_stride<float, true, true, true, false>(B, c);
}
}
