#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 5 -c 1 -f -o gpurun_out/prof_r02_plain10 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --robots 296 --no-cpu-baseline --no-configs --chunk-robots 296 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 9 -c 1 -f -o gpurun_out/prof_r02_collide6 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 6 --robots 888 --selfcol --sparse 0.55 --no-cpu-baseline --no-configs --chunk-robots 888 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 5 -c 1 -f -o gpurun_out/prof_r02_fluid8 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 8 --robots 296 --workload water --no-cpu-baseline --no-configs --chunk-robots 296 > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log | cut -c1-100
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_under_ncu_r02.log 2>&1
ls -la gpurun_out/prof_r02_* gpurun_out/launches_r02.csv
