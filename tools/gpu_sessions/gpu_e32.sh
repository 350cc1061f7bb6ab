#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_evaluate_all.py tests/test_abi.py -m gpu -x -q > gpurun_out/pytest_e32.log 2>&1; tail -3 gpurun_out/pytest_e32.log | cut -c1-300
timeout 900 python tools/ga_generation.py --robots 1024 --resident 8 --ref-sample 32 --out gpurun_out/ga_generation_r02c.json > gpurun_out/ga_e32.log 2>&1; grep -E "voxelyze:" gpurun_out/ga_e32.log | cut -c100-230; grep -A10 "wall_s_per_generation" gpurun_out/ga_e32.log | head -12; grep speedup gpurun_out/ga_e32.log
