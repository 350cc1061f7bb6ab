#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shards_a_population" > gpurun_out/pytest_e34.log 2>&1; tail -12 gpurun_out/pytest_e34.log | cut -c1-400
