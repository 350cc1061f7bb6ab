#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_e11_full.log 2>&1; tail -6 gpurun_out/pytest_e11_full.log | cut -c1-250
timeout 600 python tools/ga_generation.py --robots 1024 --ref-sample 0 > gpurun_out/ga_e11.log 2>&1; grep -E "voxelyze:|wall_s|results" gpurun_out/ga_e11.log
timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-configs > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('10^3 value %.4e frac %.3f e2e %.4e'%(d['value'], d['roofline']['frac'], d['e2e']['value']))" || tail -3 gpurun_out/b.err
# ncu full captures: plain 10^3 (actuated), self-collision on sparse 6^3, fluid 8^3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 40 -c 1 -f -o gpurun_out/prof_r02_plain10 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --robots 296 --no-cpu-baseline --no-configs --chunk-robots 296 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 40 -c 1 -f -o gpurun_out/prof_r02_collide6 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 6 --robots 888 --selfcol --sparse 0.55 --no-cpu-baseline --no-configs --chunk-robots 888 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 40 -c 1 -f -o gpurun_out/prof_r02_fluid8 \
    python bench.py --steps 3 --warmup 3 --timesteps 1000 --lattice 8 --robots 296 --workload water --no-cpu-baseline --no-configs --chunk-robots 296 > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log | cut -c1-100
ls -la gpurun_out/prof_r02_*
