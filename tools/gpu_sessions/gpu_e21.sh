#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "largest_supported" > gpurun_out/pytest_e21.log 2>&1; tail -15 gpurun_out/pytest_e21.log | cut -c1-300
