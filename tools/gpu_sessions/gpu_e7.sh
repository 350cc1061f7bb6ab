#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 5 -c 1 -f -o gpurun_out/prof_e7 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --robots 296 --no-cpu-baseline > gpurun_out/ncu_e7.log 2>&1
tail -2 gpurun_out/ncu_e7.log | cut -c1-200
