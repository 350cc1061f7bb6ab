#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 5 -c 1 -f -o gpurun_out/prof_r02b_col10 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 10 --robots 296 --chunk-robots 296 --selfcol --no-cpu-baseline --no-configs > gpurun_out/ncu_e39.log 2>&1
tail -n 2 gpurun_out/ncu_e39.log | cut -c1-160
