#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 5 -c 1 -f -o gpurun_out/prof_r02_sparse6col \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 6 --robots 1184 --chunk-robots 1184 --selfcol --sparse 0.55 --no-cpu-baseline --no-configs > gpurun_out/ncu_e23.log 2>&1
tail -2 gpurun_out/ncu_e23.log | cut -c1-200
