#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_evaluate_all.py -m gpu -x -q > gpurun_out/pytest_e27.log 2>&1; tail -4 gpurun_out/pytest_e27.log | cut -c1-300
timeout 900 python tools/ga_generation.py --robots 1024 --resident 6 --ref-sample 32 --out gpurun_out/ga_generation_r02b.json > gpurun_out/ga_e27.log 2>&1; grep -E "voxelyze:|wall_s|results|speedup|wall_s_per_generation" -A3 gpurun_out/ga_e27.log | head -60
