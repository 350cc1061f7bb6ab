#!/bin/bash
# ncu capture of the cluster form (lib_v2: relaxed arrive)
mkdir -p gpurun_out
VXG_LIB=$PWD/evosoro_b200/lib_v2/libvxgpu.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_c2 -s 5 -c 1 -f -o gpurun_out/prof_r02_cluster10 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --robots 296 --chunk-robots 296 --no-cpu-baseline --no-configs > gpurun_out/ncu_e20.log 2>&1
tail -2 gpurun_out/ncu_e20.log | cut -c1-200
