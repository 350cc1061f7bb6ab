#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ga_generation.py --robots 1024 --ref-sample 0 > gpurun_out/ga_e18.log 2>&1; grep -E "voxelyze:|vxg_pop_create|wall_s|results" gpurun_out/ga_e18.log
timeout 600 python tools/ga_generation.py --robots 1024 --ref-sample 0 > gpurun_out/ga_e18b.log 2>&1; grep -E "voxelyze:|vxg_pop_create|wall_s|results" gpurun_out/ga_e18b.log
timeout 300 python -m pytest tests/test_population.py -m gpu -q 2>&1 | tail -2
