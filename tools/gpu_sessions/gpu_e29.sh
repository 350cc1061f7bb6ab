#!/bin/bash
# end-of-round ncu captures: self-collision kernel on the CPPN-like population, land/water kernel at 8^3 (two robots per SM), launch list of the bench
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 5 -c 1 -f -o gpurun_out/prof_r02b_sparse6col \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 6 --robots 1184 --chunk-robots 1184 --selfcol --sparse 0.55 --no-cpu-baseline --no-configs > gpurun_out/ncu_e29a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 5 -c 1 -f -o gpurun_out/prof_r02b_fluid8 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 8 --robots 592 --chunk-robots 592 --workload water --no-cpu-baseline --no-configs > gpurun_out/ncu_e29b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_under_ncu_r02b.log 2>&1
tail -1 gpurun_out/ncu_e29a.log gpurun_out/ncu_e29b.log | cut -c1-160; wc -l gpurun_out/launches_r02b.csv
